"""Shared test inputs: seeded graphs, weights, and tolerant comparisons."""
import numpy as np
from scipy import sparse

RTOL = 1e-5  # north star: embeddings and scores within 1e-5 relative (fp32)


def random_bipartite(n_users, n_items, n_pos, seed=0, n_props=0, n_links=0, dup_links=0):
    """Symmetrised COO adjacency in the reference's entry order (+ optional property links with duplicates)."""
    rng = np.random.RandomState(seed)
    keys = np.unique(rng.randint(0, n_users * n_items, size=n_pos))
    rng.shuffle(keys)
    r, c = keys // n_items, keys % n_items + n_users
    n = n_users + n_items + n_props
    if n_props:
        ii = rng.randint(0, n_items, size=n_links) + n_users
        pp = rng.randint(0, n_props, size=n_links) + n_users + n_items
        pick = rng.randint(0, n_links, size=dup_links)
        r = np.concatenate([r, ii, ii[pick]])
        c = np.concatenate([c, pp, pp[pick]])
    rows = np.concatenate([r, c]).astype(np.int32)
    cols = np.concatenate([c, r]).astype(np.int32)
    return sparse.coo_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=(n, n), dtype=np.float32)


def glorot(rng, shape):
    lim = np.sqrt(6.0 / (shape[-2] + shape[-1])) if len(shape) > 1 else np.sqrt(3.0 / shape[0])
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def assert_close(got, want, rtol=RTOL, atol=None, what=""):
    """max |got-want| <= rtol * max|want| (relative to the tensor scale) unless atol is given."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, "{}: shape {} vs {}".format(what, got.shape, want.shape)
    scale = max(np.abs(want).max(), 1e-30) if want.size else 1.0
    tol = atol if atol is not None else rtol * scale
    err = np.abs(got - want).max() if want.size else 0.0
    assert err <= tol, "{}: max abs err {:.3e} > {:.3e} (scale {:.3e})".format(what, err, tol, scale)


def assert_topk_equivalent(ids, vals, oracle_scores, k, tol=2e-6):
    """ids must be bit-identical to the oracle's stable top-k wherever the oracle's score gaps
    exceed `tol`; inside a near-tie (fp32 exp differs by an ulp between CPU and GPU) the returned
    item must still carry an oracle score within tol of the oracle's item at that rank."""
    s = np.asarray(oracle_scores, np.float64)
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    ref_vals = np.take_along_axis(s, order, axis=1)
    ids = np.asarray(ids)
    got_vals = np.take_along_axis(s, ids.astype(np.int64), axis=1)
    assert np.abs(got_vals - ref_vals).max() <= tol, "top-k scores differ beyond tolerance"
    assert np.abs(np.asarray(vals, np.float64) - got_vals).max() <= max(tol, 1e-5), "returned scores differ from the oracle's"
    sorted_s = -np.sort(-s, axis=1)[:, :k + 1]
    if sorted_s.shape[1] < k + 1:  # k == catalog size: nothing ranks below the last slot
        sorted_s = np.concatenate([sorted_s, np.full((len(s), 1), -np.inf)], axis=1)
    gaps = sorted_s[:, :-1] - sorted_s[:, 1:]
    prev_gap = np.concatenate([np.full((len(s), 1), np.inf), gaps[:, :-1]], axis=1)
    clear = (gaps > tol) & (prev_gap > tol)
    assert (ids[clear] == order[clear]).all(), "top-k ids differ where the ranking is unambiguous"
    return float(clear.mean()) if clear.size else 1.0


def weights_struct(w, proj_first=None):
    """{product weight path: array} -> the numpy structures the oracle takes (Keras layouts: kernel [in,out]).
    proj_first: {fusion layer name: bool} for attention fusions that project."""
    w = dict(w)

    def stack(prefix):
        ks = sorted({int(n[len(prefix):].split("/")[0]) for n in w if n.startswith(prefix)})
        return [(w["%s%d/kernel" % (prefix, k)], w["%s%d/bias" % (prefix, k)]) for k in ks]

    out = dict(embeddings=w["gnn/gnn_layers/embeddings"], layers=[])
    k = 0
    while any(n.startswith("gnn/gnn_layers/seq_layers.%d/" % k) for n in w):
        pre = "gnn/gnn_layers/seq_layers.%d/" % k
        lw = {n[len(pre):]: v for n, v in w.items() if n.startswith(pre)}
        if "kernel" in lw and lw["kernel"].ndim == 3:  # GATConv stores [F,1,H]
            lw["kernel"] = lw["kernel"].reshape(lw["kernel"].shape[0], -1)
        if "attn_kernel_self" in lw:
            lw["attn_self"] = lw.pop("attn_kernel_self").reshape(-1)
            lw["attn_neigh"] = lw.pop("attn_kernel_neigh").reshape(-1)
        out["layers"].append(lw)
        k += 1
    if any(n.startswith("rs/unet/") or n.startswith("rs/inet/") for n in w) or not any(n.startswith("rs/dense1a/") for n in w):
        out.update(unet=stack("rs/unet/layers."), inet=stack("rs/inet/layers."), clf=stack("rs/clf/layers."))
    else:
        for name in ("dense1a", "dense1b", "dense2a", "dense2b", "dense3a", "dense3b", "clf"):
            out[name] = stack("rs/%s/layers." % name)
        if any(n.startswith("rs/residual/") for n in w):
            out["residual"] = stack("rs/residual/layers.")
        for name in ("fuse1a", "fuse1b", "fuse2"):
            if "rs/%s/att_weight" % name in w:
                out[name] = dict(att_weight=w["rs/%s/att_weight" % name], proj_weight=w.get("rs/%s/proj_weight" % name),
                                 proj_first=(proj_first or {}).get(name))
    return out


def export_weights(model):
    """Model weights as the numpy structures the oracle takes."""
    named = {n: w.detach().cpu().numpy() for n, w in model.named_weights()}
    pf = {name: getattr(model.rs, name).proj_first for name in ("fuse1a", "fuse1b", "fuse2")
          if hasattr(model.rs, name) and getattr(getattr(model.rs, name), "proj_first", None) is not None}
    out = weights_struct(named, pf)
    n_layers = len(model.gnn.gnn_layers.seq_layers)
    out["layers"] += [{} for _ in range(n_layers - len(out["layers"]))]  # weight-free LightGCN layers
    return out


def kg_graphs(n_users=40, n_items=30, n_props=25, n_pos=420, n_links=150, dup_links=12):
    """Seeded (user-item [U+I], item-property [I+P]) adjacencies in the layout of type_adjacency='unary-kg'
    (preprocess.py:113-145): both symmetrised, the second with `dup_links` repeated (item, property) links - the
    same pair under two predicates - which GCN-type layers sum and GraphSage / GAT keep as separate edges."""
    ui = random_bipartite(n_users, n_items, n_pos, seed=13)
    ip = random_bipartite(n_items, n_props, n_links, seed=5)
    half = ip.nnz // 2
    pick = np.random.RandomState(9).randint(0, half, size=dup_links)
    r = np.concatenate([ip.row[:half], ip.row[pick]])
    c = np.concatenate([ip.col[:half], ip.col[pick]])
    rows, cols = np.concatenate([r, c]).astype(np.int32), np.concatenate([c, r]).astype(np.int32)
    ip = sparse.coo_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=ip.shape, dtype=np.float32)
    return ui, ip


KG_GRAPHS = {"default": dict(), "sparse": dict(n_pos=120, n_links=40, dup_links=4)}  # tests/golden/make_golden_models_kg.py


KG_COMMON = dict(n_layers=2, embedding_dim=8, l2_regularizer=1e-4, aggregate="mean", dropout_rate=0.0,
                 final_node="concatenation", activation="relu")


def kg_model(case, graphs, n_users, n_items):
    """the constructor call of tests/golden/make_golden_models_kg.py on the product's classes"""
    from deep_cbrs_amar_renaissance_b200.models import basic, hybrid
    name = case.split("-")[0]
    kw = dict(KG_COMMON, n_hiddens=[8, 8])
    if name.startswith("Hybrid"):
        kw.update(dense_units=[[12, 12], [14, 6], [16, 16]], clf_units=[16, 16], feature_based=True)
    else:
        kw.update(dense_units=[12, 12], clf_units=[16, 16])
    if case.endswith("-itemconcat"):
        kw["item_node"] = "concatenation"
    if case.endswith("-uiconcat"):
        kw["user_item_node"] = "concatenation"
    if case.endswith("-mean"):
        kw.update(final_node="mean", aggregate="sum")
    adjs = graphs if "TW" in name else graphs[:2]
    return getattr(hybrid if name.startswith("Hybrid") else basic, name)(n_users, n_items, adjs, **kw), kw


def relation_blocks(a_hat, first_rel1_node):
    """The normalised adjacency split into the two relations of SURVEY row R by node range: an entry belongs to
    relation 1 when it touches a node >= first_rel1_node (item <-> property links), else to relation 0; self loops ride
    on relation 0 (the device build's self_rel).  Returns [A_0, A_1] as scipy CSR."""
    coo = a_hat.tocoo()
    is_r1 = (coo.row >= first_rel1_node) | (coo.col >= first_rel1_node)
    is_r1 = np.where(coo.row == coo.col, False, is_r1)
    return [sparse.csr_matrix((coo.data[pick], (coo.row[pick], coo.col[pick])), shape=coo.shape) for pick in (~is_r1, is_r1)]
