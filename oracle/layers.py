"""Oracle (TEST INFRASTRUCTURE): layer arithmetic, rows P0-P5, S1-S4, T, R.

float32 throughout.  Sparse products accumulate each output row sequentially
in ascending column order (scipy csr_matvecs walks the stored order; TF-CPU
sparse_dense_matmul does the same on its row-major reordered tensor).
"""
import numpy as np
from scipy import sparse

F32 = np.float32


def _csr(indptr, indices, data, n_cols=None):
    n = len(indptr) - 1
    return sparse.csr_matrix((np.asarray(data, F32), np.asarray(indices), np.asarray(indptr)),
                             shape=(n, n if n_cols is None else n_cols))


def relu(x):
    return np.maximum(x, F32(0))


def sigmoid(x):
    x = np.asarray(x, F32)
    return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def activation(name):
    if name is None or name == "linear":
        return lambda x: x
    if name == "relu":
        return relu
    if name == "sigmoid":
        return sigmoid
    if name == "tanh":
        return lambda x: np.tanh(x, dtype=F32)
    raise ValueError("activation not restated: {}".format(name))


# ----------------------------------------------------------------- Keras Dense
def dense(x, kernel, bias=None, act=None):
    """[3P] keras.layers.Dense: act(x @ kernel[in,out] + bias)."""
    y = np.asarray(x, F32) @ np.asarray(kernel, F32)
    if bias is not None:
        y = y + np.asarray(bias, F32)
    return activation(act)(y.astype(F32))


def dense_network(x, layers, act="relu"):
    """models/dense.py:4-7 - Sequential of Dense(u, activation)."""
    for k, b in layers:
        x = dense(x, k, b, act)
    return x


def dense_classifier(x, layers, act="relu"):
    """models/dense.py:10-17 - hidden Dense(u, act) then Dense(1, sigmoid)."""
    for k, b in layers[:-1]:
        x = dense(x, k, b, act)
    k, b = layers[-1]
    return dense(x, k, b, "sigmoid")


def bf16_round(x):
    """float32 -> nearest bfloat16 (ties to even) -> float32: what storing a tensor as bf16 does to it"""
    u = np.ascontiguousarray(x, dtype=F32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(F32)


# ------------------------------------------------------------------------- P1
def gcn_conv(x, a_hat, kernel, bias, act="relu", operand_dtype="fp32"):
    """[3P] spektral GCNConv.call: act(A_hat @ (x @ kernel) + bias). Transform first.
    operand_dtype='bf16' restates the product's optional bf16 STORAGE of the transform (products and sums fp32)."""
    z = (np.asarray(x, F32) @ np.asarray(kernel, F32)).astype(F32)
    if operand_dtype == "bf16":
        z = bf16_round(z)
    y = (a_hat @ z).astype(F32)
    if bias is not None:
        y = y + np.asarray(bias, F32)
    return activation(act)(y)


# ------------------------------------------------------------------------- P4
def lightgcn_conv(x, a_hat):
    """layers/lightgcn_conv.py:51-54: X' = A_hat X."""
    return (a_hat @ np.asarray(x, F32)).astype(F32)


# ------------------------------------------------------------------------- (f)-3 DGCF
DGCF_EPSILONS = [1e-1, 1e-2, 1e-3, 5e-4]


def dgcf_preprocess(adj_coo):
    """layers/dgcf_conv.py:38-80 (the reference's own scipy code, restated): crosshop = a.dot(a);
    a, crosshop = gcn_filter(a), gcn_filter(crosshop); crosshop = high_pass_filter(a, crosshop)
    (keep entries > eps for the eps in [1e-1, 1e-2, 1e-3, 5e-4] whose surviving count is closest in ratio to
    a's edge count, first on ties); return a + crosshop + eye.  float32 CSR, sorted indices.
    Returns (matrix, info dict)."""
    from .graph import gcn_filter
    a = sparse.csr_matrix(adj_coo, dtype=F32)
    a.sum_duplicates()
    crosshop = a.dot(a)
    a_hat, cross_hat = gcn_filter(adj_coo), gcn_filter(crosshop.tocoo())
    edges = len(a_hat.data)
    filtered = [cross_hat.multiply(cross_hat > F32(eps)).tocsr() for eps in DGCF_EPSILONS]
    for f in filtered:
        f.eliminate_zeros()
    cross_edges = [len(f.data) for f in filtered]
    ratios = [float("inf") if c == 0 else (edges / c if edges > c else c / edges) for c in cross_edges]
    best = int(np.argmin(ratios))
    out = (a_hat + filtered[best] + sparse.eye(a_hat.shape[0], dtype=F32, format="csr")).tocsr().astype(F32)
    out.sort_indices()
    return out, dict(edges=edges, cross_edges=cross_edges, ratios=ratios, epsilon=DGCF_EPSILONS[best])


def dgcf_conv(x, m, w):
    """DGCFConv.call (dgcf_conv.py:32-36) with LocalityAdaptive (:101-102): M (x * sigmoid(w)), w [N,1]."""
    gate = sigmoid(np.asarray(w, F32).reshape(-1, 1))
    return (m @ (np.asarray(x, F32) * gate).astype(F32)).astype(F32)


# ------------------------------------------------------------------------- P2
def sage_aggregate(x, indptr, indices, aggregate="mean"):
    """Neighbourhood aggregate over the RAW edge list (values ignored, dups kept)."""
    n = len(indptr) - 1
    ones = np.ones(len(indices), F32)
    a = _csr(indptr, indices, ones)
    s = (a @ np.asarray(x, F32)).astype(F32)
    if aggregate == "sum":
        return s
    if aggregate == "mean":
        cnt = np.diff(indptr).astype(F32)
        out = np.zeros_like(s)
        nz = cnt > 0
        out[nz] = s[nz] / cnt[nz, None]
        return out
    if aggregate in ("max", "min"):
        x = np.asarray(x, F32)
        out = np.zeros((n, x.shape[1]), F32)
        red = np.maximum if aggregate == "max" else np.minimum
        for i in range(n):
            nb = indices[indptr[i]:indptr[i + 1]]
            if len(nb):
                out[i] = red.reduce(x[nb], axis=0)
        return out
    raise ValueError("aggregate not restated: {}".format(aggregate))


def sage_conv(x, indptr, indices, kernel, bias, aggregate="mean", act="relu"):
    """[3P] spektral GraphSageConv: act(l2_normalize([x || agg] @ kernel + bias)).

    l2_normalize(v) = v * rsqrt(max(sum(v^2), 1e-12)); normalise BEFORE activation.
    """
    x = np.asarray(x, F32)
    agg = sage_aggregate(x, indptr, indices, aggregate)
    o = (np.concatenate([x, agg], axis=1) @ np.asarray(kernel, F32)).astype(F32)
    if bias is not None:
        o = o + np.asarray(bias, F32)
    ss = np.maximum((o * o).sum(axis=1, keepdims=True, dtype=F32), F32(1e-12))
    o = (o / np.sqrt(ss, dtype=F32)).astype(F32)
    return activation(act)(o)


# ------------------------------------------------------------------------- P3
def gat_conv(x, indptr, indices, kernel, attn_self, attn_neigh, bias, act="relu",
             add_self_loops=True):
    """[3P] spektral GATConv (1 head, dropout 0), sparse single mode.

    kernel [F,H] (Keras stores [F,1,H]); attn_self / attn_neigh [H].
    """
    from .graph import gat_edges
    x = np.asarray(x, F32)
    n = x.shape[0]
    z = (x @ np.asarray(kernel, F32).reshape(x.shape[1], -1)).astype(F32)
    p = (z * np.asarray(attn_self, F32).reshape(1, -1)).sum(axis=1, dtype=F32)
    q = (z * np.asarray(attn_neigh, F32).reshape(1, -1)).sum(axis=1, dtype=F32)
    if add_self_loops:
        ptr, cols = gat_edges(indptr, indices)
    else:
        ptr, cols = np.asarray(indptr), np.asarray(indices)
    rows = np.repeat(np.arange(n), np.diff(ptr))
    e = (p[rows] + q[cols]).astype(F32)
    e = np.where(e > 0, e, F32(0.2) * e).astype(F32)
    m = np.full(n, -np.inf, F32)
    np.maximum.at(m, rows, e)
    w = np.exp(e - m[rows], dtype=F32)
    s = _csr(ptr, cols, w) @ np.ones((n, 1), F32)
    alpha = (w / (s[rows, 0] + F32(1e-9))).astype(F32)
    out = (_csr(ptr, cols, alpha) @ z).astype(F32)
    if bias is not None:
        out = out + np.asarray(bias, F32)
    return activation(act)(out)


# ------------------------------------------------------------------------- P5
def reduce_layers(hs, method="concatenation", w=None):
    """layers/reduction.py:15-33 (+ WeightedSum :54-55: sum(w^2 * h))."""
    if method == "concatenation":
        return np.concatenate(hs, axis=-1)
    if method == "sum":
        acc = hs[0].astype(F32)
        for h in hs[1:]:
            acc = acc + h
        return acc
    if method == "mean":
        acc = hs[0].astype(F32)
        for h in hs[1:]:
            acc = acc + h
        return (acc / F32(len(hs))).astype(F32)
    if method == "last":
        return hs[-1]
    if method == "w-sum":
        w = np.ones(len(hs), F32) if w is None else np.asarray(w, F32).ravel()
        acc = (w[0] * w[0]) * hs[0]
        for wi, h in zip(w[1:], hs[1:]):
            acc = acc + (wi * wi) * h
        return acc.astype(F32)
    raise ValueError("Reduction method not supported: " + method)


# ------------------------------------------------------------------------- P0
def propagate(kind, emb, graph, layer_weights, final_node="concatenation", aggregate="mean", operand_dtype="fp32"):
    """SequentialGNN.call, models/gnn.py:74-84: x=E; hs=[x]; for l: x=layer([x,A]); reduce(hs).

    kind in {'gcn','lightgcn','sage','gat'}; graph = scipy CSR A_hat for
    gcn/lightgcn, (indptr, indices) of the raw reordered adjacency otherwise.
    """
    x = np.asarray(emb, F32)
    hs = [x]
    for w in layer_weights:
        if kind == "gcn":
            x = gcn_conv(x, graph, w["kernel"], w["bias"], operand_dtype=operand_dtype)
        elif kind == "lightgcn":
            x = lightgcn_conv(x, graph)
        elif kind == "sage":
            x = sage_conv(x, graph[0], graph[1], w["kernel"], w["bias"], aggregate)
        elif kind == "gat":
            x = gat_conv(x, graph[0], graph[1], w["kernel"], w["attn_self"], w["attn_neigh"], w["bias"])
        elif kind == "dgcf":
            x = dgcf_conv(x, graph, w["locality_adaptive/locality-adaptive-weights"])
        else:
            raise ValueError(kind)
        hs.append(x)
    if kind in ("lightgcn", "dgcf"):
        final_node = "mean"  # gnn.py:378
    return reduce_layers(hs, final_node)


# ------------------------------------------------------------------------- (f)-4 Two-Step / Two-Way
def two_step(kind, step_one, step_two, n_items, item_node="mean", final_node="concatenation", aggregate="mean"):
    """TwoStepGNN.call, models/tsgnn.py:92-94 with HalfInputSequentialGNN.call, models/gnn.py:136-147:
    x = SequentialGNN_kg(None) reduced by item_node; second input = concat([user embeddings, x[:n_items]], axis 0);
    the second loop runs over the user-item graph and is reduced by final_node.
    step_one / step_two = dict(embeddings=, graph=, layers=[...])."""
    x = propagate(kind, step_one["embeddings"], step_one["graph"], step_one["layers"], item_node, aggregate)
    if kind in ("lightgcn", "dgcf") and item_node != "mean":
        raise ValueError("propagate() forces 'mean' for weight-free families; item_node must be 'mean' there")
    x2 = np.concatenate([np.asarray(step_two["embeddings"], F32), x[:n_items]], axis=0)
    return propagate(kind, x2, step_two["graph"], step_two["layers"], final_node, aggregate)


def two_way(kind, way_one, way_two, step_two, n_users, n_items, user_item_node="mean", final_node="concatenation",
            aggregate="mean"):
    """TwoWayGNN.call, models/twgnn.py:93-100 with FullInputSequentialGNN.call, models/gnn.py:198-207:
    users = SequentialGNN_up(None)[:n_users]; items = SequentialGNN_ip(None)[:n_items]; the stacked rows are the whole
    input of the third loop over the user-item graph (no embeddings of its own)."""
    users = propagate(kind, way_one["embeddings"], way_one["graph"], way_one["layers"], user_item_node, aggregate)
    items = propagate(kind, way_two["embeddings"], way_two["graph"], way_two["layers"], user_item_node, aggregate)
    x = np.concatenate([users[:n_users], items[:n_items]], axis=0)
    return propagate(kind, x, step_two["graph"], step_two["layers"], final_node, aggregate)


# ------------------------------------------------------------------------- R
def rgcn_conv(x, rel_graphs, kernels, bias, act="relu"):
    """Relational extension (no reference counterpart).

    out = act(sum_r A_hat_r @ (x @ W_r) + b).  With one relation this is gcn_conv.
    Relations are summed in ascending r; inside a relation, ascending column.
    """
    acc = None
    for a_r, w_r in zip(rel_graphs, kernels):
        t = (a_r @ (np.asarray(x, F32) @ np.asarray(w_r, F32)).astype(F32)).astype(F32)
        acc = t if acc is None else acc + t
    if bias is not None:
        acc = acc + np.asarray(bias, F32)
    return activation(act)(acc)


# -------------------------------------------------------------------- S1 + S2
def basic_rs(emb, u_ids, i_ids, unet, inet, clf, act="relu"):
    """BasicGNN.embed_recommend + BasicRS.call, models/basic.py:31-37,65-75."""
    u = dense_network(emb[u_ids], unet, act)
    i = dense_network(emb[i_ids], inet, act)
    return dense_classifier(np.concatenate([u, i], axis=1), clf, act)


# -------------------------------------------------------------------- S3 + S4
def fusion(a, b, fw=None):
    """layers/fusion.py:49-68.  fw None -> concatenate; else dict(att_weight [, proj_weight, proj_first])."""
    if fw is None:
        return np.concatenate([a, b], axis=1)
    if fw.get("proj_weight") is not None:
        if fw["proj_first"]:
            a = (a @ np.asarray(fw["proj_weight"], F32)).astype(F32)
        else:
            b = (b @ np.asarray(fw["proj_weight"], F32)).astype(F32)
    x = np.stack([a, b], axis=1).astype(F32)
    att = np.tanh(x @ np.asarray(fw["att_weight"], F32), dtype=F32)
    e = np.exp(att - att.max(axis=1, keepdims=True), dtype=F32)
    att = e / e.sum(axis=1, keepdims=True)
    return (att * x).sum(axis=1).astype(F32)


def hybrid_cbrs(emb, u_ids, i_ids, u_bert, i_bert, w, act="relu", feature_based=True):
    """HybridBertGNN.embed_recommend + HybridCBRS.call, models/hybrid.py:72-89,130-140.

    w = dict(dense1a, dense1b, dense2a, dense2b, dense3a, dense3b, clf), each a
    list of (kernel, bias).  concatenate fusion only (layers/fusion.py:51-53).
    """
    ug = dense_network(emb[u_ids], w["dense1a"], act)
    ig = dense_network(emb[i_ids], w["dense1b"], act)
    ub = dense_network(np.asarray(u_bert, F32), w["dense2a"], act)
    ib = dense_network(np.asarray(i_bert, F32), w["dense2b"], act)
    if feature_based:
        x1 = dense_network(fusion(ug, ig, w.get("fuse1a")), w["dense3a"], act)
        x2 = dense_network(fusion(ub, ib, w.get("fuse1b")), w["dense3b"], act)
    else:
        x1 = dense_network(fusion(ug, ub, w.get("fuse1a")), w["dense3a"], act)
        x2 = dense_network(fusion(ig, ib, w.get("fuse1b")), w["dense3b"], act)
    x = fusion(x1, x2, w.get("fuse2"))
    if w.get("residual") is None:
        return dense_classifier(x, w["clf"], act)
    # models/hybrid.py:89 + models/dense.py:20-27: last residual Dense is linear, activation after the add
    r = dense_network(x, w["residual"][:-1], act)
    k, b = w["residual"][-1]
    r = dense(r, k, b, None)
    return dense_classifier(activation(act)(r + x1 + x2), w["clf"], act)


# ------------------------------------------------------------------------- T
def top_k_pairs(u_ids, i_ids, scores, k):
    """Per-user top-k among an explicit pair list.

    Restates utilities/metrics.py:21-34: stable sort by (user asc, score desc)
    then head(k) per user; ties keep input order (pandas multi-key sort is
    stable).  Users are returned in ascending order (the reference iterates a
    Python set, whose order is unspecified).
    Returns (users [M], items [M], scores [M], input_rows [M]).
    """
    u_ids = np.asarray(u_ids)
    scores = np.asarray(scores).ravel()
    order = np.lexsort((np.arange(len(scores)), -scores.astype(np.float64), u_ids))
    su = u_ids[order]
    start = np.r_[0, np.flatnonzero(su[1:] != su[:-1]) + 1]
    rank = np.arange(len(su)) - np.repeat(start, np.diff(np.r_[start, len(su)]))
    sel = order[rank < k]
    return u_ids[sel], np.asarray(i_ids)[sel], scores[sel], sel


def top_k_catalog(score_matrix, k):
    """Full-catalog top-k (new capability): reference scorer on every
    (u,i) + stable descending sort => ties go to the lower item index.
    Returns (ids int32 [U,k], scores float32 [U,k])."""
    s = np.asarray(score_matrix, F32)
    order = np.argsort(-s.astype(np.float64), axis=1, kind="stable")[:, :k]
    return order.astype(np.int32), np.take_along_axis(s, order, axis=1)
