"""Golden vectors from the REFERENCE's own Two-Step / Two-Way model code (scope row (f)-4), run unmodified end to end.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_models_kg.py

src/models/{tsgnn,twgnn,gnn,basic,hybrid}.py and src/data/preprocess.py are imported UNMODIFIED from
/root/reference/src with oracle/tf_np_stub on sys.path, exactly as in make_golden_models.py (same leaves, same caveats:
GCNConv / GraphSageConv / GATConv / gcn_filter are the oracle's restatement, LightGCNConv / DGCFConv the reference's own
code).  What these vectors pin is the reference's own wiring of the variants: SequentialGNN -> slice of the item (and
user) rows -> HalfInputSequentialGNN / FullInputSequentialGNN (gnn.py:87-207), the width bookkeeping of
tsgnn.py:66-77 / twgnn.py:70-77, item_node / user_item_node, and the user-property graph of get_user_properties.
Output (committed): tests/golden/models/golden_models_kg.npz; keys follow the product's weight paths.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle", "tf_np_stub"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, REPO)
if not hasattr(np, "mat"):          # numpy 2 removed np.mat (utilities/math.py:49); it was np.asmatrix
    np.mat = np.asmatrix

from data.preprocess import get_user_properties  # noqa: E402  (the reference's modules)
from models import basic, hybrid  # noqa: E402
from tests.helpers import kg_graphs  # noqa: E402

N_USERS, N_ITEMS, N_PROPS, BATCH, BERT = 40, 30, 25, 48, 10
BASIC = dict(dense_units=[12, 12], clf_units=[16, 16])
HYBRID = dict(dense_units=[[12, 12], [14, 6], [16, 16]], clf_units=[16, 16], feature_based=True)
CASES = {
    "BasicTSGCN": BASIC, "BasicTSGraphSage": BASIC, "BasicTSGAT": BASIC, "BasicTSLightGCN": BASIC, "BasicTSDGCF": BASIC,
    "BasicTSGCN-itemconcat": dict(BASIC, item_node="concatenation"),
    "BasicTSGraphSage-mean": dict(BASIC, final_node="mean", aggregate="sum"),
    "BasicTWGCN": BASIC, "BasicTWGraphSage": BASIC, "BasicTWGAT": BASIC, "BasicTWLightGCN": BASIC,
    "BasicTWGCN-uiconcat": dict(BASIC, user_item_node="concatenation"),
    "BasicTWDGCF-sparse": BASIC,
    "HybridBertTSGCN": HYBRID, "HybridBertTWGraphSage": HYBRID,
}


# '-sparse' cases run on a thinner graph: on the default one the user-property graph is so dense that no entry of
# its normalised square exceeds 0.1, and the reference's high_pass_filter divides by that zero count
# (dgcf_conv.py:75) - BasicTWDGCF cannot be built there at all.
GRAPHS = {"default": dict(), "sparse": dict(n_pos=120, n_links=40, dup_links=4)}


def common():
    """a FRESH dict per model: the reference extends the n_hiddens list it is given in place (tsgnn.py:70-72)"""
    return dict(n_hiddens=[8, 8], n_layers=2, embedding_dim=8, l2_regularizer=1e-4, aggregate="mean", dropout_rate=0.0,
                final_node="concatenation", activation="relu")


def stack_weights(prefix, seq, out):
    for k, layer in enumerate(seq.layers):
        out["%s/layers.%d/kernel" % (prefix, k)] = layer.kernel
        out["%s/layers.%d/bias" % (prefix, k)] = layer.bias


def collect(model):
    w = {}
    for part in ("step_one_gnn_layers", "way_one_gnn_layers", "way_two_gnn_layers", "step_two_gnn_layers"):
        seq = getattr(model.gnn, part, None)
        if seq is None:
            continue
        if getattr(seq, "embeddings", None) is not None:
            w["gnn/%s/embeddings" % part] = seq.embeddings
        for l, layer in enumerate(seq.seq_layers):
            pre = "gnn/%s/seq_layers.%d/" % (part, l)
            for name, arr in layer.weights.items():
                w[pre + name] = arr
            if hasattr(layer, "locality_adaptive"):
                w[pre + "locality_adaptive/locality-adaptive-weights"] = layer.locality_adaptive.w
    rs = model.rs
    for name in ("unet", "inet", "clf", "dense1a", "dense1b", "dense2a", "dense2b", "dense3a", "dense3b"):
        if hasattr(rs, name):
            stack_weights("rs/" + name, getattr(rs, name), w)
    return w


def main():
    rng = np.random.RandomState(78)
    g = dict(n_users=np.array(N_USERS), n_items=np.array(N_ITEMS), n_props=np.array(N_PROPS))
    graphs = {}
    for tag, gkw in GRAPHS.items():
        ui, ip = kg_graphs(N_USERS, N_ITEMS, N_PROPS, **gkw)
        up = get_user_properties(ui, ip, N_USERS, N_ITEMS)   # the reference's own (dense detour)
        graphs[tag] = (ui, ip, up)
        g.update({tag + "/up_row": up.row, tag + "/up_col": up.col, tag + "/up_data": up.data})
    u = rng.randint(0, N_USERS, size=BATCH)
    i = rng.randint(0, N_ITEMS, size=BATCH) + N_USERS
    ub = rng.standard_normal((BATCH, BERT)).astype(np.float32)
    ib = rng.standard_normal((BATCH, BERT)).astype(np.float32)
    g.update(u=u, i=i, ub=ub, ib=ib)
    for case, extra in CASES.items():
        cls_name = case.split("-")[0]
        is_hybrid = cls_name.startswith("Hybrid")
        cls = getattr(hybrid if is_hybrid else basic, cls_name)
        kw = common()
        kw.update({k: (list(v) if isinstance(v, list) else v) for k, v in extra.items()})
        ui, ip, up = graphs["sparse" if case.endswith("-sparse") else "default"]
        adjs = (ui, ip, up) if "TW" in cls_name else (ui, ip)
        model = cls(N_USERS, N_ITEMS, adjs, **kw)
        scores = model((u, i, ub, ib)) if is_hybrid else model((u, i))
        emb = model.gnn(None)
        g[case + "/out/embeddings"] = np.asarray(emb, np.float32)
        g[case + "/out/scores"] = np.asarray(scores, np.float32)
        g[case + "/out/n_hiddens"] = np.asarray(getattr(model.gnn, "n_hiddens", []), np.int64)
        for name, arr in collect(model).items():
            g[case + "/" + name] = np.asarray(arr)
        print(case, "emb", np.asarray(emb).shape, "scores %.4f..%.4f" % (float(scores.min()), float(scores.max())),
              "n_hiddens", getattr(model.gnn, "n_hiddens", None))
    np.savez_compressed(os.path.join(HERE, "models", "golden_models_kg.npz"), **g)
    print("wrote", len(g), "arrays")


if __name__ == "__main__":
    main()
