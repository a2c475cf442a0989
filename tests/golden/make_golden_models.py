"""Golden vectors from the REFERENCE's own MODEL code, run unmodified end to end.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_models.py

src/models/{basic,hybrid,gnn,dense}.py and src/layers/* are imported UNMODIFIED from /root/reference/src with
oracle/tf_np_stub on sys.path.  What runs is the reference's own wiring: SequentialGNN's layer loop and reduction
(gnn.py:74-84), the per-family builders (gnn.py:267-415), the embedding lookups (basic.py:65-75, hybrid.py:130-140),
BasicRS and HybridCBRS in every mode the grids use (feature/entity based, attention fusion, residual classifier),
the dense builders.  The LEAVES are stand-ins ([3P], oracle/tf_np_stub/README.md): keras Dense by its definition,
tf.nn.embedding_lookup as indexing, LightGCNConv/DGCFConv through the reference's own code over modal_dot = A @ x,
and GCNConv / GraphSageConv / GATConv / gcn_filter through the ORACLE's restatement - so these vectors pin the wiring
for all families and the full arithmetic only where the layer is the reference's own (LightGCN, DGCF).
Output (committed): tests/golden/models/golden_models.npz; keys follow the product's weight paths.
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

from models import basic, hybrid  # noqa: E402  (the reference's modules)
from tests.helpers import random_bipartite  # noqa: E402

N_USERS, N_ITEMS, BATCH, BERT = 40, 30, 48, 10
COMMON = dict(n_hiddens=[8, 8], n_layers=2, embedding_dim=8, l2_regularizer=1e-4, aggregate="mean", dropout_rate=0.0,
              final_node="concatenation", activation="relu")
CASES = {
    "BasicGCN": ("basic", dict(dense_units=[12, 12], clf_units=[16, 16])),
    "BasicGAT": ("basic", dict(dense_units=[12, 12], clf_units=[16, 16])),
    "BasicGraphSage": ("basic", dict(dense_units=[12, 12], clf_units=[16, 16])),
    "BasicLightGCN": ("basic", dict(dense_units=[12, 12], clf_units=[16, 16])),
    "BasicDGCF": ("basic", dict(dense_units=[12, 12], clf_units=[16, 16])),
    "BasicGCN-notowers": ("basic", dict(dense_units=[], clf_units=[16])),
    "HybridBertGCN-feature": ("hybrid", dict(dense_units=[[12, 12], [14, 6], [16, 16]], clf_units=[16, 16], feature_based=True)),
    "HybridBertGCN-entity": ("hybrid", dict(dense_units=[[12, 12], [14, 6], [16, 16]], clf_units=[16, 16], feature_based=False)),
    "HybridBertGAT-attention": ("hybrid", dict(dense_units=[[12, 12], [14, 6], [16, 16]], clf_units=[16, 16], feature_based=True,
                                               fusion_method="attention")),
    "HybridBertLightGCN-residual": ("hybrid", dict(dense_units=[[12, 12], [14, 6], [16, 16]], clf_units=[16, 16],
                                                   feature_based=True, residual=True)),
    "HybridBertGraphSage-entity-attention": ("hybrid", dict(dense_units=[[12, 12], [14, 6], [16, 16]], clf_units=[16, 16],
                                                           feature_based=False, fusion_method="attention")),
}


def stack_weights(prefix, seq, out):
    for k, layer in enumerate(seq.layers):
        out["%s/layers.%d/kernel" % (prefix, k)] = layer.kernel
        out["%s/layers.%d/bias" % (prefix, k)] = layer.bias


def collect(model, kind):
    w = {"gnn/gnn_layers/embeddings": model.gnn.gnn_layers.embeddings}
    for l, layer in enumerate(model.gnn.gnn_layers.seq_layers):
        pre = "gnn/gnn_layers/seq_layers.%d/" % l
        for name, arr in layer.weights.items():
            w[pre + name] = arr
        if hasattr(layer, "locality_adaptive"):
            w[pre + "locality_adaptive/locality-adaptive-weights"] = layer.locality_adaptive.w
    rs = model.rs
    if kind == "basic":
        for name in ("unet", "inet", "clf"):
            stack_weights("rs/" + name, getattr(rs, name), w)
    else:
        for name in ("dense1a", "dense1b", "dense2a", "dense2b", "dense3a", "dense3b", "clf"):
            stack_weights("rs/" + name, getattr(rs, name), w)
        if rs.residual is not None:
            stack_weights("rs/residual", rs.residual, w)
        for name in ("fuse1a", "fuse1b", "fuse2"):
            f = getattr(rs, name)
            if f.method == "attention":
                w["rs/%s/att_weight" % name] = f.att_weight
                if f.proj_first is not None:
                    w["rs/%s/proj_weight" % name] = f.proj_weight
                    w["rs/%s/proj_first" % name] = np.array(bool(f.proj_first))
    return w


def main():
    rng = np.random.RandomState(77)
    adj = random_bipartite(N_USERS, N_ITEMS, 420, seed=13)
    g = dict(adj_row=adj.row, adj_col=adj.col, adj_data=adj.data, n_nodes=np.array(adj.shape[0]))
    u = rng.randint(0, N_USERS, size=BATCH)
    u[:6] = u[0]
    i = rng.randint(0, N_ITEMS, size=BATCH) + N_USERS
    ub = rng.standard_normal((BATCH, BERT)).astype(np.float32)
    ib = rng.standard_normal((BATCH, BERT)).astype(np.float32)
    g.update(u=u, i=i, ub=ub, ib=ib)
    for case, (kind, extra) in CASES.items():
        cls_name = case.split("-")[0]
        cls = getattr(basic if kind == "basic" else hybrid, cls_name)
        kw = dict(COMMON)
        kw.update(extra)
        model = cls(adj, **kw)
        scores = model((u, i)) if kind == "basic" else model((u, i, ub, ib))
        emb = model.gnn(None)
        g[case + "/out/embeddings"] = np.asarray(emb, np.float32)
        g[case + "/out/scores"] = np.asarray(scores, np.float32)
        for name, arr in collect(model, kind).items():
            g[case + "/" + name] = np.asarray(arr)
        print(case, "emb", np.asarray(emb).shape, "scores %.4f..%.4f" % (float(scores.min()), float(scores.max())))
    os.makedirs(os.path.join(HERE, "models"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "models", "golden_models.npz"), **g)
    print("wrote", len(g), "arrays")


if __name__ == "__main__":
    main()
