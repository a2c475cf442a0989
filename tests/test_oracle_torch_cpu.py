"""CPU: the timed torch-CPU baseline computes what the numpy oracle computes; the numpy twin of
the device graph generator is deterministic and well formed."""
import numpy as np
import torch
from scipy import sparse

from oracle import graph as og
from oracle import layers as ol
from oracle import synth as osynth
from oracle import torch_cpu as oc
from tests.helpers import assert_close


def test_torch_cpu_baseline_matches_numpy_oracle():
    n_users, n_items = 400, 60
    row, col = osynth.synth_bipartite(n_users, n_items, 5000, 42)
    n = n_users + n_items
    a_t, nnz = oc.gcn_filter_torch(row, col, n)
    adj = sparse.coo_matrix((np.ones(len(row), np.float32), (row, col)), shape=(n, n))
    a_np = og.gcn_filter(adj)
    assert nnz == a_np.nnz
    rng = np.random.RandomState(0)
    emb = rng.standard_normal((n, 16)).astype(np.float32)
    layers = [(oc.glorot(rng, (16, 16)), rng.standard_normal(16).astype(np.float32) * 0.1) for _ in range(2)]
    got = oc.gcn_forward(torch.from_numpy(emb), a_t, [(torch.from_numpy(w), torch.from_numpy(b)) for w, b in layers])
    want = ol.propagate("gcn", emb, a_np, [dict(kernel=w, bias=b) for w, b in layers])
    assert_close(got.numpy(), want)
    mlp = oc.random_basic_rs(rng, 48, [24, 24], [16, 16])
    u = rng.randint(0, n_users, size=100)
    i = rng.randint(0, n_items, size=100) + n_users
    s = oc.basic_rs(got, torch.from_numpy(u), torch.from_numpy(i), mlp)
    npw = {k: [(a.numpy(), b.numpy()) for a, b in v] for k, v in mlp.items()}
    assert_close(s.numpy(), ol.basic_rs(want, u, i, npw["unet"], npw["inet"], npw["clf"]))


def test_synthetic_graph_shape():
    n_users, n_items, n_edges = 3000, 2000, 40000
    row, col = osynth.synth_bipartite(n_users, n_items, n_edges, 7)
    assert row.dtype == np.int32 and len(row) == 2 * n_edges
    assert (row[:n_edges] < n_users).all() and (col[:n_edges] >= n_users).all() and (col[:n_edges] < n_users + n_items).all()
    assert np.array_equal(row[n_edges:], col[:n_edges]) and np.array_equal(col[n_edges:], row[:n_edges])
    r2, c2 = osynth.synth_bipartite(n_users, n_items, n_edges, 7)
    assert np.array_equal(row, r2) and np.array_equal(col, c2)
    pop = np.bincount(col[:n_edges] - n_users, minlength=n_items)
    assert pop.max() > 8 * max(np.median(pop), 1)  # heavy-tailed item popularity
