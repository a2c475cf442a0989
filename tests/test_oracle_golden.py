"""CPU: the oracle and the host-side data code against fixtures produced by the
REFERENCE's own loaders (tests/golden/make_golden.py).  Rows G0, G1, S0, S3."""
import os

import numpy as np
import pytest

from deep_cbrs_amar_renaissance_b200.data import loaders
from oracle import graph as og

CASES = ["ui_small", "uip_small", "hybrid_small"]


def _load(golden_dir, case):
    root = os.path.join(golden_dir, case)
    return root, np.load(os.path.join(root, "golden.npz"))


def _tsv(path):
    return np.loadtxt(path, dtype=np.int64, delimiter="\t", ndmin=2)


@pytest.mark.parametrize("case", CASES)
def test_oracle_id_compaction_and_adjacency(golden_dir, case):
    root, g = _load(golden_dir, case)
    train, test, users, items = og.compact_ids(_tsv(os.path.join(root, "train2id.tsv")),
                                               _tsv(os.path.join(root, "test2id.tsv")))
    assert np.array_equal(train, g["train_ratings"]) and np.array_equal(test, g["test_ratings"])
    assert np.array_equal(users, g["users"]) and np.array_equal(items, g["items"])
    props_path = os.path.join(root, "props2id.tsv")
    if os.path.exists(props_path):
        triples, props = og.compact_props(_tsv(props_path), items)
        adj = og.build_adjacency(train, len(users), len(items), triples, len(props), "unary-uip")
    else:
        adj = og.build_adjacency(train, len(users), len(items))
    assert tuple(adj.shape) == tuple(g["adj_shape"])
    assert adj.row.dtype == np.int32 and adj.data.dtype == np.float32
    assert np.array_equal(adj.row, g["adj_row"]) and np.array_equal(adj.col, g["adj_col"])
    assert np.array_equal(adj.data, g["adj_data"])
    # duplicate summing == scipy tocsr of the reference's matrix
    csr = adj.tocsr()
    assert np.array_equal(csr.indptr, g["csr_indptr"]) and np.array_equal(csr.indices, g["csr_indices"])
    assert np.array_equal(csr.data, g["csr_data"])


@pytest.mark.parametrize("case", CASES)
def test_host_loaders_match_reference(golden_dir, case):
    root, g = _load(golden_dir, case)
    kw = dict(train_ratings_filepath=os.path.join(root, "train2id.tsv"),
              test_ratings_filepath=os.path.join(root, "test2id.tsv"), train_batch_size=128, test_batch_size=64)
    if os.path.exists(os.path.join(root, "props2id.tsv")):
        kw.update(props_triples_filepath=os.path.join(root, "props2id.tsv"), type_adjacency="unary-uip")
    hybrid = os.path.exists(os.path.join(root, "user-lastlayer.json"))
    if hybrid:
        kw.update(bert_user_filepath=os.path.join(root, "user-lastlayer.json"),
                  bert_item_filepath=os.path.join(root, "item-lastlayer.json"))
        train, test = loaders.load_user_item_graph_bert_embeddings(**kw)
    else:
        train, test = loaders.load_user_item_graph(**kw)
    assert np.array_equal(train.ratings, g["train_ratings"]) and np.array_equal(test.ratings, g["test_ratings"])
    assert np.array_equal(train.users, g["users"]) and np.array_equal(train.items, g["items"])
    adj = train.adj_matrix
    assert np.array_equal(adj.row, g["adj_row"]) and np.array_equal(adj.col, g["adj_col"])
    assert np.array_equal(adj.data, g["adj_data"]) and adj.dtype == np.float32
    assert len(train) == int(g["n_train_batches"]) and len(test) == int(g["n_test_batches"])
    # batch contents over two epochs: the gather indices of the hot path
    for ep in range(2):
        for b in range(len(train)):
            x, y = train[b]
            assert np.array_equal(x[0], g["train_ep%d_b%d_u" % (ep, b)])
            assert np.array_equal(x[1], g["train_ep%d_b%d_i" % (ep, b)])
            assert np.array_equal(y, g["train_ep%d_b%d_y" % (ep, b)])
            if hybrid:
                assert np.array_equal(x[2], g["train_ep%d_b%d_ub" % (ep, b)])
                assert np.array_equal(x[3], g["train_ep%d_b%d_ib" % (ep, b)])
        train.on_epoch_end()
    for b in range(len(test)):
        x, y = test[b]
        assert np.array_equal(x[0], g["test_ep0_b%d_u" % b]) and np.array_equal(x[1], g["test_ep0_b%d_i" % b])


@pytest.mark.parametrize("sym", [True, False])
def test_kg_graphs_match_reference(golden_dir, sym):
    """'unary-kg' + user_properties=True (Two-Step / Two-Way inputs): the three adjacencies from the reference's own
    loader (tests/golden/make_golden_kg.py) vs the oracle's restatement and the product's all-sparse host code."""
    root = os.path.join(golden_dir, "uip_small")
    g = np.load(os.path.join(root, "golden_kg.npz"))
    tag = "sym" if sym else "dir"
    train, _, users, items = og.compact_ids(_tsv(os.path.join(root, "train2id.tsv")), _tsv(os.path.join(root, "test2id.tsv")))
    triples, props = og.compact_props(_tsv(os.path.join(root, "props2id.tsv")), items)
    ui, ip = og.build_kg_adjacencies(train, len(users), len(items), triples, len(props), symmetric=sym)
    oracle = (ui, ip, og.user_properties(ui, ip, len(users), len(items)))
    trainset, _ = loaders.load_user_item_graph(
        os.path.join(root, "train2id.tsv"), os.path.join(root, "test2id.tsv"), os.path.join(root, "props2id.tsv"),
        type_adjacency="unary-kg", user_properties=True, symmetric_adjacency=sym)
    assert len(trainset.adj_matrix) == 3
    for mats in (oracle, trainset.adj_matrix):
        for name, m in zip(("ui", "ip", "up"), mats):
            assert tuple(m.shape) == tuple(g["%s_%s_shape" % (tag, name)])
            for field in ("row", "col", "data"):
                want = g["%s_%s_%s" % (tag, name, field)]
                got = getattr(m, field)
                assert got.dtype == want.dtype and np.array_equal(got, want), (name, field)
    # without user_properties the loader hands over the pair (loaders.py:319)
    pair, _ = loaders.load_user_item_graph(
        os.path.join(root, "train2id.tsv"), os.path.join(root, "test2id.tsv"), os.path.join(root, "props2id.tsv"),
        type_adjacency="unary-kg", symmetric_adjacency=sym)
    assert len(pair.adj_matrix) == 2 and np.array_equal(pair.adj_matrix[1].row, g[tag + "_ip_row"])


def test_user_properties_on_a_larger_graph_stays_sparse():
    """the all-sparse product == the oracle's dense restatement on a seeded graph with duplicate item-property links"""
    from deep_cbrs_amar_renaissance_b200.data.preprocess import get_user_properties
    from tests.helpers import random_bipartite
    n_users, n_items, n_props = 120, 70, 45
    ui = random_bipartite(n_users, n_items, 900, seed=3)
    ip = random_bipartite(n_items, n_props, 260, seed=4)
    dup = np.arange(0, 40)
    ip = type(ip)((np.concatenate([ip.data, ip.data[dup]]), (np.concatenate([ip.row, ip.row[dup]]),
                                                              np.concatenate([ip.col, ip.col[dup]]))), shape=ip.shape)
    want = og.user_properties(ui, ip, n_users, n_items)
    got = get_user_properties(ui, ip, n_users, n_items)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert np.array_equal(got.row, want.row) and np.array_equal(got.col, want.col) and np.array_equal(got.data, want.data)


def test_unknown_test_id_raises(golden_dir, tmp_path):
    root, _ = _load(golden_dir, "ui_small")
    test = _tsv(os.path.join(root, "test2id.tsv")).copy()
    test[0, 0] = 10 ** 9
    bad = tmp_path / "test2id.tsv"
    np.savetxt(bad, test, fmt="%d", delimiter="\t")
    with pytest.raises(KeyError):
        loaders.load_user_item_graph(os.path.join(root, "train2id.tsv"), str(bad))
    with pytest.raises(KeyError):
        og.compact_ids(_tsv(os.path.join(root, "train2id.tsv")), test)


def test_unknown_adjacency_type_raises(golden_dir):
    root, _ = _load(golden_dir, "ui_small")
    with pytest.raises(ValueError):
        loaders.load_user_item_graph(os.path.join(root, "train2id.tsv"), os.path.join(root, "test2id.tsv"),
                                     type_adjacency="nope")


def test_embedding_loaders_match_reference(golden_dir):
    """load_graph_embeddings / load_bert_embeddings / load_hybrid_embeddings (the pre-computed-embedding baselines,
    loaders.py:147-271) against batches produced by the reference's own loaders (tests/golden/make_golden_kge.py)"""
    root = os.path.join(golden_dir, "hybrid_small")
    g = np.load(os.path.join(root, "golden_kge.npz"))
    base = dict(train_ratings_filepath=os.path.join(root, "train2id.tsv"), test_ratings_filepath=os.path.join(root, "test2id.tsv"),
                train_batch_size=128, test_batch_size=64)
    bert = dict(bert_user_filepath=os.path.join(root, "user-lastlayer.json"), bert_item_filepath=os.path.join(root, "item-lastlayer.json"))
    graph = dict(graph_filepath=os.path.join(root, "768TransH.json"))
    for name, fn, kw, width in (("graph", loaders.load_graph_embeddings, graph, 2), ("bert", loaders.load_bert_embeddings, bert, 2),
                                ("hybrid", loaders.load_hybrid_embeddings, dict(graph, **bert), 4)):
        tr, te = fn(**base, **kw)
        assert [len(tr), len(te)] == list(g[name + "_n_batches"])
        for tag, seq, epochs in (("train", tr, 2), ("test", te, 1)):
            for ep in range(epochs):
                for b in range(len(seq)):
                    x, y = seq[b]
                    assert len(x) == width
                    for k, arr in enumerate(x):
                        want = g["%s_%s_ep%d_b%d_x%d" % (name, tag, ep, b, k)]
                        assert arr.dtype == want.dtype and np.array_equal(arr, want)
                    assert np.array_equal(y, g["%s_%s_ep%d_b%d_y" % (name, tag, ep, b)])
                seq.on_epoch_end()
