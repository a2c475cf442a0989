"""CPU: pins for the restated Spektral / Keras arithmetic.  The reference ships no tests or
golden tensors for it (parity unpinned, see oracle/__init__.py); these are the algebraic
self-checks SURVEY.md 8c lists plus cross-checks of the scipy fast path against plain loops."""
import numpy as np
from scipy import sparse

from oracle import graph as og
from oracle import layers as ol
from tests.helpers import assert_close, glorot, random_bipartite


def test_inv_sqrt_matches_numpy_power():
    deg = np.arange(1, 200001, dtype=np.float32)
    ours = og.inv_sqrt_degree(deg)
    theirs = np.power(deg, np.float32(-0.5))
    ulp = np.abs(ours.view(np.int32).astype(np.int64) - theirs.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1  # numpy's float32 pow is allowed 1 ulp; ours is the correctly rounded value
    assert og.inv_sqrt_degree(np.array([0.0], np.float32))[0] == 0.0


def test_gcn_filter_matches_scipy_formulation():
    adj = random_bipartite(30, 20, 200, seed=1, n_props=10, n_links=40, dup_links=8)
    a, b = og.gcn_filter(adj), og.gcn_filter_scipy(adj)
    assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    assert a.data.dtype == np.float32
    assert_close(a.data, b.data, rtol=3e-7)
    assert a.diagonal().min() > 0  # self loops present everywhere
    assert_close(a.toarray(), a.toarray().T, rtol=1e-7)  # symmetric in, symmetric out


def test_reorder_raw_keeps_duplicates_and_sorts():
    adj = random_bipartite(12, 9, 60, seed=2, n_props=5, n_links=20, dup_links=6)
    ptr, idx, val = og.reorder_raw(adj)
    assert ptr[-1] == adj.nnz == len(idx)
    for i in range(adj.shape[0]):
        seg = idx[ptr[i]:ptr[i + 1]]
        assert (np.diff(seg) >= 0).all()
    assert adj.tocsr().nnz < adj.nnz  # there really are duplicates


def test_spmm_order_matches_python_loop():
    rng = np.random.RandomState(0)
    a = og.gcn_filter(random_bipartite(8, 6, 30, seed=3))
    x = rng.standard_normal((a.shape[0], 5)).astype(np.float32)
    want = np.zeros_like(x)
    for i in range(a.shape[0]):
        acc = np.zeros(5, np.float32)
        for j in range(a.indptr[i], a.indptr[i + 1]):
            acc = (acc + a.data[j] * x[a.indices[j]]).astype(np.float32)
        want[i] = acc
    assert np.array_equal(ol.lightgcn_conv(x, a), want)


def test_lightgcn_on_regular_graph_is_averaging():
    n, k = 12, 4  # circulant k-regular graph: A_hat = (A + I) / (k + 1)
    rows = np.repeat(np.arange(n), k)
    cols = (rows + np.tile([1, 2, n - 1, n - 2], n)) % n
    adj = sparse.coo_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=(n, n))
    x = np.random.RandomState(1).standard_normal((n, 3)).astype(np.float32)
    want = (adj.toarray() + np.eye(n)) @ x / (k + 1)
    assert_close(ol.lightgcn_conv(x, og.gcn_filter(adj)), want, rtol=1e-6)


def test_gcn_with_identity_weight_is_relu_lightgcn():
    adj = random_bipartite(20, 15, 120, seed=4)
    a = og.gcn_filter(adj)
    x = np.random.RandomState(2).standard_normal((35, 6)).astype(np.float32)
    got = ol.gcn_conv(x, a, np.eye(6, dtype=np.float32), np.zeros(6, np.float32))
    assert_close(got, np.maximum(ol.lightgcn_conv(x, a), 0), rtol=1e-6)


def test_gat_with_zero_attention_is_mean_over_closed_neighbourhood():
    adj = random_bipartite(15, 10, 70, seed=5)
    ptr, idx, _ = og.reorder_raw(adj)
    rng = np.random.RandomState(3)
    x = rng.standard_normal((25, 4)).astype(np.float32)
    w = glorot(rng, (4, 7))
    zeros = np.zeros(7, np.float32)
    got = ol.gat_conv(x, ptr, idx, w, zeros, zeros, None, act=None)
    dense = (adj.toarray() > 0).astype(np.float64)  # dedup: this graph has none
    np.fill_diagonal(dense, 1.0)
    want = (dense / dense.sum(1, keepdims=True)) @ (x @ w)
    assert_close(got, want, rtol=1e-5)


def test_gat_counts_duplicate_edges_twice():
    rows = np.array([0, 0, 1, 2], np.int32)
    cols = np.array([1, 1, 0, 0], np.int32)  # (0,1) twice
    adj = sparse.coo_matrix((np.ones(4, np.float32), (rows, cols)), shape=(3, 3))
    ptr, idx, _ = og.reorder_raw(adj)
    x = np.eye(3, dtype=np.float32)
    got = ol.gat_conv(x, ptr, idx, np.eye(3, dtype=np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32),
                      None, act=None)
    assert_close(got[0], np.array([1 / 3, 2 / 3, 0.0]), rtol=1e-6)  # self + 2 x neighbour 1
    agg = ol.sage_aggregate(x, ptr, idx, "mean")
    assert_close(agg[0], np.array([0.0, 1.0, 0.0]), rtol=1e-6)
    assert np.array_equal(agg[1], x[0]) and np.array_equal(ol.sage_aggregate(x, ptr, idx, "sum")[0], 2 * x[1])


def test_sage_normalises_before_relu_and_handles_isolated_nodes():
    adj = sparse.coo_matrix((np.ones(2, np.float32), ([0, 1], [1, 0])), shape=(3, 3))  # node 2 isolated
    ptr, idx, _ = og.reorder_raw(adj)
    rng = np.random.RandomState(4)
    x = rng.standard_normal((3, 4)).astype(np.float32)
    w = glorot(rng, (8, 5))
    b = rng.standard_normal(5).astype(np.float32)
    got = ol.sage_conv(x, ptr, idx, w, b)
    pre = np.concatenate([x, ol.sage_aggregate(x, ptr, idx)], 1) @ w + b
    want = np.maximum(pre / np.linalg.norm(pre, axis=1, keepdims=True), 0)
    assert_close(got, want, rtol=1e-6)
    assert np.array_equal(ol.sage_aggregate(x, ptr, idx)[2], np.zeros(4, np.float32))
    zero_row = ol.sage_conv(np.zeros((3, 4), np.float32), ptr, idx, w, np.zeros(5, np.float32))
    assert np.array_equal(zero_row, np.zeros((3, 5), np.float32))  # 1e-12 floor, no NaN


def test_rgcn_with_one_relation_is_gcn():
    adj = random_bipartite(10, 8, 50, seed=6)
    a = og.gcn_filter(adj)
    rng = np.random.RandomState(5)
    x = rng.standard_normal((18, 4)).astype(np.float32)
    w, b = glorot(rng, (4, 6)), rng.standard_normal(6).astype(np.float32)
    assert np.array_equal(ol.rgcn_conv(x, [a], [w], b), ol.gcn_conv(x, a, w, b))


def test_reductions():
    hs = [np.full((2, 3), v, np.float32) for v in (1.0, 2.0, 6.0)]
    assert ol.reduce_layers(hs, "concatenation").shape == (2, 9)
    assert np.array_equal(ol.reduce_layers(hs, "sum"), np.full((2, 3), 9.0, np.float32))
    assert np.array_equal(ol.reduce_layers(hs, "mean"), np.full((2, 3), 3.0, np.float32))
    assert np.array_equal(ol.reduce_layers(hs, "last"), hs[-1])
    assert np.array_equal(ol.reduce_layers(hs, "w-sum", [1, 2, 0.5]), np.full((2, 3), 1 + 8 + 1.5, np.float32))


def test_top_k_pairs_is_stable_and_per_user():
    u = np.array([1, 0, 1, 1, 0, 1])
    i = np.array([10, 11, 12, 13, 14, 15])
    s = np.array([0.5, 0.9, 0.7, 0.7, 0.9, 0.1], np.float32)
    uu, ii, ss, rows = ol.top_k_pairs(u, i, s, 2)
    assert uu.tolist() == [0, 0, 1, 1]
    assert ii.tolist() == [11, 14, 12, 13]  # ties keep input order
    ids, vals = ol.top_k_catalog(np.array([[0.5, 0.9, 0.9, 0.1]], np.float32), 3)
    assert ids.tolist() == [[1, 2, 0]]


def test_known_parameter_counts_from_the_report():
    """doc.pdf Table 17 (N = 6,036 + 3,192 = 9,228): pins weight shapes of the restated layers."""
    n, d, h = 9228, 16, 16
    rs = lambda din, du, c: 2 * (din * du + du + du * du + du) + (2 * du * c + c) + (c * c + c) + (c + 1)  # noqa: E731
    gcn = n * d + 2 * (d * h + h) + rs(d + 2 * h, 48, 64)
    sage = n * d + 2 * (2 * d * h + h) + rs(d + 2 * h, 48, 64)
    light = n * d + rs(d, 48, 64)
    assert gcn == 168033 and sage == 168545 and light == 164417
    assert 9228 * 8 + 2 * (8 * 8 + 8) + 2 * (24 * 24 + 24 + 24 * 24 + 24) + (48 * 48 + 48) + (48 * 48 + 48) + 49 == 81121
