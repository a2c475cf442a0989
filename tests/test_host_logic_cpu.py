"""Host-side logic that needs no GPU: optimiser configuration, weight shapes of the hybrid tweaks, builders."""
import math

import numpy as np
import pytest

from deep_cbrs_amar_renaissance_b200 import training
from deep_cbrs_amar_renaissance_b200.layers import FusionLayer
from deep_cbrs_amar_renaissance_b200.models.dense import (build_dense_classifier, build_dense_network,
                                                          build_residual_dense_network)
from deep_cbrs_amar_renaissance_b200.models.hybrid import HybridCBRS


def test_adam_from_config_forms():
    a = training.Adam.from_config({"learning_rate": 1e-2, "beta_1": 0.8, "name": "Adam"})
    assert (a.learning_rate, a.beta_1, a.beta_2, a.epsilon) == (1e-2, 0.8, 0.999, 1e-7)
    assert training.Adam.from_config(None).learning_rate == 1e-3 and training.Adam.from_config("adam").beta_1 == 0.9
    assert training.Adam.from_config(a) is a

    class KerasLike:  # what optimizers.Adam(**config) looks like to the caller (experiment.py:158)
        learning_rate, beta_1, beta_2, epsilon = 5e-4, 0.9, 0.99, 1e-8

    b = training.Adam.from_config(KerasLike())
    assert (b.learning_rate, b.beta_2, b.epsilon) == (5e-4, 0.99, 1e-8)
    with pytest.raises(NotImplementedError):
        training.Adam.from_config("sgd")


def test_adam_rate_is_the_keras_bias_correction():
    a = training.Adam(learning_rate=1e-3)
    for t in (1, 2, 10, 1000):
        assert math.isclose(a.rate(t), 1e-3 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t), rel_tol=1e-12)
    assert math.isclose(a.rate(1), 1e-3 * math.sqrt(0.001) / 0.1, rel_tol=1e-12)


def test_fusion_layer_weights_follow_the_reference_shapes():
    f = FusionLayer('attention')
    assert f.build_for(64, 64) == 64 and tuple(f.att_weight.shape) == (64, 64) and f.proj_weight is None
    g = FusionLayer('attention')   # narrower first input is projected up (fusion.py:24-32)
    assert g.build_for(32, 48) == 48 and g.proj_first is True and tuple(g.proj_weight.shape) == (32, 48)
    h = FusionLayer('attention')
    assert h.build_for(48, 32) == 48 and h.proj_first is False and tuple(h.proj_weight.shape) == (32, 48)
    c = FusionLayer('concatenate')
    assert c.build_for(10, 7) == 17 and c.weights == []
    with pytest.raises(ValueError):
        FusionLayer('sum')


def test_dense_builders():
    net = build_dense_network([48, 48], activation='relu')
    assert [l.activation for l in net.layers] == ['relu', 'relu']
    clf = build_dense_classifier([64, 64], n_classes=1, activation='relu')
    assert [(l.units, l.activation) for l in clf.layers] == [(64, 'relu'), (64, 'relu'), (1, 'sigmoid')]
    res = build_residual_dense_network([64, 64], activation='relu')
    assert [(l.units, l.activation) for l in res.layers] == [(64, 'relu'), (64, None)]   # dense.py:20-27
    assert [(l.units, l.activation) for l in build_dense_classifier([], n_classes=1).layers] == [(1, 'sigmoid')]


def test_hybrid_cbrs_parameter_counts_for_the_tweaks():
    units = [[48, 48], [256, 64], [64, 64]]
    plain = HybridCBRS(feature_based=True, dense_units=units, clf_units=[64, 64])
    plain.build_for(48, 768)
    att = HybridCBRS(feature_based=True, dense_units=units, clf_units=[64, 64], fusion_method='attention')
    att.build_for(48, 768)
    res = HybridCBRS(feature_based=True, dense_units=units, clf_units=[64, 64], residual=True)
    res.build_for(48, 768)
    n_plain = plain.count_params()
    # attention fuses the two 64-wide branches into ONE 64-wide vector: + 64*64 attention weights, and the
    # classifier's first kernel shrinks from 128x64 to 64x64
    assert att.count_params() == n_plain + 64 * 64 - 64 * 64
    # residual: the [64, 64] classifier body becomes the residual stack (same shapes); the head reads 64 features
    assert res.count_params() == n_plain
    with pytest.raises(ValueError):
        HybridCBRS(dense_units=[[48], [64], [32]], clf_units=[64], residual=True)


# ---------------------------------------------------------------- Two-Step / Two-Way wiring (scope row (f)-4)
class _ShapeOnlyGraph:
    def __init__(self, adj):
        self.shape = tuple(adj.shape)
        self.n_nodes = self.shape[0]


@pytest.fixture
def shape_only_graphs(monkeypatch):
    """no GPU here: the adjacency upload is replaced by a shape holder, so constructors and build_weights() run"""
    from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph
    monkeypatch.setattr(DeviceGraph, "from_scipy", classmethod(lambda cls, adj, device=None, chunk_edges=None:
                                                               adj if isinstance(adj, _ShapeOnlyGraph) else _ShapeOnlyGraph(adj)))


def _kg_cases():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "models", "golden_models_kg.npz"))
    return g, sorted({k.split("/")[0] for k in g.files if "/out/" in k})


def test_two_step_two_way_weights_follow_the_reference_run(shape_only_graphs):
    """same weight paths and shapes as the reference's own models built with the same keywords (goldens), and the
    n_hiddens bookkeeping of tsgnn.py:66-77 / twgnn.py:70-77 - made on a copy, the caller's list is left alone"""
    from deep_cbrs_amar_renaissance_b200.data.preprocess import get_user_properties
    from tests.helpers import kg_graphs, kg_model
    g, cases = _kg_cases()
    n_users, n_items, n_props = int(g["n_users"]), int(g["n_items"]), int(g["n_props"])
    ui, ip = kg_graphs(n_users, n_items, n_props)
    graphs = (ui, ip, get_user_properties(ui, ip, n_users, n_items))
    for case in cases:
        model, kw = kg_model(case, graphs, n_users, n_items)
        model.build_weights(10) if case.startswith("Hybrid") else model.build_weights()
        got = {nm: tuple(w.shape) for nm, w in model.named_weights()}
        want = {k[len(case) + 1:]: tuple(g[k].shape) for k in g.files if k.startswith(case + "/") and "/out/" not in k}
        assert got == want, (case, sorted(set(got.items()) ^ set(want.items())))
        assert list(getattr(model.gnn, "n_hiddens", [])) == list(g[case + "/out/n_hiddens"])
        assert kw["n_hiddens"] == [8, 8]
        assert model.gnn.out_dim == g[case + "/out/embeddings"].shape[1]


def test_two_step_two_way_constructor_errors(shape_only_graphs):
    from deep_cbrs_amar_renaissance_b200.models import basic
    from tests.helpers import kg_graphs
    ui, ip = kg_graphs()
    with pytest.raises(ValueError, match="two adjacency"):
        basic.BasicTSGCN(40, 30, (ui,), n_hiddens=[8, 8])
    with pytest.raises(ValueError, match="three adjacency"):
        basic.BasicTWGCN(40, 30, (ui, ip), n_hiddens=[8, 8])
    with pytest.raises(NotImplementedError):
        basic.BasicTSGCN(40, 30, (ui, ip), n_hiddens=[8, 8], cache_neighbours=True)
    # experiment.py:139-146 dispatches on these parents
    assert issubclass(basic.BasicTSGAT, basic.BasicTSGNN) and issubclass(basic.BasicTWDGCF, basic.BasicTWGNN)
    assert issubclass(basic.BasicTSGNN, basic.BasicGNN)


def test_typed_edges_from_the_loader_follow_the_entry_order_of_the_reference_graph():
    """relations= (row R, extension): the COO entries are exactly the reference's untyped 'unary-uip' graph (golden
    fixture written by the reference's own loader), each with a relation id; both directions of an edge agree."""
    import os
    from deep_cbrs_amar_renaissance_b200.data import loaders
    from deep_cbrs_amar_renaissance_b200.data.preprocess import edge_relations
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "uip_small")
    files = [os.path.join(root, f) for f in ("train2id.tsv", "test2id.tsv", "props2id.tsv")]
    gold = np.load(os.path.join(root, "golden.npz"))
    plain, _ = loaders.load_user_item_graph(*files, type_adjacency="unary-uip")
    assert np.array_equal(plain.adj_matrix.row, gold["adj_row"]) and np.array_equal(plain.adj_matrix.col, gold["adj_col"])
    raw = np.loadtxt(files[2], dtype=np.int64, delimiter="\t")
    n_users, n_items = len(plain.users), len(plain.items)
    for mode in ("node-range", "predicate"):
        typed, _ = loaders.load_user_item_graph(*files, type_adjacency="unary-uip", relations=mode)
        adj = typed.adj_matrix
        assert np.array_equal(adj.coo.row, gold["adj_row"]) and np.array_equal(adj.coo.col, gold["adj_col"])
        assert np.array_equal(adj.coo.data, gold["adj_data"]) and adj.shape == tuple(gold["adj_shape"])
        is_prop_edge = (adj.coo.row >= n_users + n_items) | (adj.coo.col >= n_users + n_items)
        assert np.array_equal(adj.rel > 0, is_prop_edge)
        half = adj.coo.nnz // 2
        assert np.array_equal(adj.rel[:half], adj.rel[half:])
        if mode == "node-range":
            assert adj.n_rel == 2 and set(np.unique(adj.rel)) == {0, 1}
        else:
            preds = np.unique(raw[:, 2])
            assert adj.n_rel == 1 + len(preds)
            n_liked = int((typed.ratings[:, 2] == 1).sum())
            assert np.array_equal(adj.rel[n_liked:half], 1 + np.searchsorted(preds, raw[:, 2]))
        blocks = adj.relation_blocks()
        assert sum(b.nnz for b in blocks) == adj.coo.nnz
        assert abs(sum(b.tocsr() for b in blocks) - adj.coo.tocsr()).nnz == 0
    with pytest.raises(ValueError):
        loaders.load_user_item_graph(*files[:2], type_adjacency="unary", relations="node-range")
    with pytest.raises(ValueError):
        edge_relations(3, raw[:, 2], "by-colour", True)


def test_bf16_stored_sources_need_the_bf16_scorer():
    """a source kept as bf16 (set_content_table(dtype='bf16')) can only be read by the TMA-fed tensor-core kernel: the Dense
    layer says so before any kernel is called - there is no silent conversion and no CPU path"""
    import torch
    from deep_cbrs_amar_renaissance_b200 import ops
    from deep_cbrs_amar_renaissance_b200.layers.dense import Dense, set_scorer_precision
    layer = Dense(16, "relu")
    table = torch.zeros(8, 64, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="set_scorer_precision"):      # fp32 precision
        layer.call_sources([(table, None)])
    set_scorer_precision(layer, "bf16")
    odd = Dense(16, "relu")
    set_scorer_precision(odd, "bf16")
    with pytest.raises(ValueError, match="multiples of 64"):           # a width the kernel does not take
        odd.call_sources([(torch.zeros(8, 96, dtype=torch.bfloat16), None)])
    with pytest.raises(ValueError, match="set_scorer_precision"):      # mixed storage of the two sources
        layer.call_sources([(table, None), (torch.zeros(8, 64), None)])
    assert ops.dense_tc_bf16_eligible(768, 0, 256) and ops.dense_tc_bf16_eligible(64, 64, 64)
    assert not ops.dense_tc_bf16_eligible(96, 0, 16) and not ops.dense_tc_bf16_eligible(64, 0, 257)
    with pytest.raises(ValueError):
        set_scorer_precision(layer, "fp16")


def test_nvlink_counters_degrade_to_none_without_nvml():
    """bench.py's NVLink byte counters: None (reported as unavailable) wherever NVML cannot serve them - never an exception"""
    import bench
    assert bench.nvlink_counters(0) is None or len(bench.nvlink_counters(0)) == 2
