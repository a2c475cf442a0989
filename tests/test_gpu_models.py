"""GPU parity of the reference-facing model API (BasicGCN / BasicGAT / BasicGraphSage /
BasicLightGCN / HybridBert* / RGCN extension) against the oracle at fixed weights."""
import os

import numpy as np
import pytest
import torch

from oracle import graph as og
from oracle import layers as ol
from tests.helpers import assert_close, assert_topk_equivalent, export_weights, glorot, random_bipartite

pytestmark = pytest.mark.gpu

KINDS = {"BasicGCN": "gcn", "BasicGraphSage": "sage", "BasicGAT": "gat", "BasicLightGCN": "lightgcn", "BasicDGCF": "dgcf"}
# the six grids of econfigs/basic-gnn.yaml: (embedding_dim, n_hiddens, dense_units, clf_units)
GRIDS = [(8, [8, 8], [24, 24], [48, 48]), (16, [16, 16], [48, 48], [64, 64]), (32, [32, 32], [96, 48], [64, 64]),
         (8, [8, 8, 8], [32, 32], [64, 64]), (16, [16, 16, 16], [64, 64], [64, 64]), (32, [32, 32, 32], [128, 64], [64, 64])]


@pytest.fixture(scope="module", autouse=True)
def _device():
    from deep_cbrs_amar_renaissance_b200 import ops
    ops.check_device()
    torch.cuda.set_device(0)


def _oracle_graph(kind, adj):
    if kind in ("gcn", "lightgcn"):
        return og.gcn_filter(adj)
    if kind == "dgcf":
        return ol.dgcf_preprocess(adj)[0]
    ptr, idx, _ = og.reorder_raw(adj)
    return (ptr, idx)


def _randomise(model, seed=0):
    """Non-zero biases and attention vectors so every term of the formulas is exercised."""
    rng = np.random.RandomState(seed)
    ws = model.get_weights()
    model.set_weights([w if w.ndim > 1 and w.shape[0] > 512 else
                       (w + rng.standard_normal(w.shape).astype(np.float32) * 0.1) for w in ws])


def _build(name, adj, grid, module="basic", **extra):
    from deep_cbrs_amar_renaissance_b200.keras_like import set_seed
    from deep_cbrs_amar_renaissance_b200.models import basic, hybrid
    set_seed(42)
    d, hid, du, cu = grid
    cls = getattr(basic if module == "basic" else hybrid, name)
    # the reference passes the whole config.model dict: unknown keys must be swallowed
    model = cls(adj, name="basic." + name, embedding_dim=d, n_hiddens=hid, n_layers=len(hid), l2_regularizer=1e-4,
                final_node="concatenation", item_node="mean", user_item_node="mean", aggregate="mean",
                dropout_rate=0.0, dense_units=du, clf_units=cu, activation="relu", **extra)
    return model


@pytest.mark.parametrize("name", sorted(KINDS))
@pytest.mark.parametrize("grid", GRIDS)
def test_basic_models_forward_matches_oracle(name, grid):
    n_users, n_items = 300, 200
    adj = random_bipartite(n_users, n_items, 6000, seed=7)
    model = _build(name, adj, grid)
    rng = np.random.RandomState(1)
    u = rng.randint(0, n_users, size=1024)
    i = rng.randint(0, n_items, size=1024) + n_users
    model((u, i))  # builds the weights (experiment.py:166)
    _randomise(model)
    got = model((u, i)).cpu().numpy()
    w = export_weights(model)
    kind = KINDS[name]
    emb = ol.propagate(kind, w["embeddings"], _oracle_graph(kind, adj), w["layers"])
    assert_close(model.gnn(None).cpu().numpy(), emb, what=name + " embeddings")
    want = ol.basic_rs(emb, u, i, w["unet"], w["inet"], w["clf"])
    assert got.shape == (1024, 1) and got.dtype == np.float32
    assert_close(got, want, what=name + " scores")
    assert (got > 0).all() and (got < 1).all()


@pytest.mark.parametrize("name", sorted(KINDS))
def test_uip_graph_with_duplicate_links(name):
    """unary-uip graphs carry duplicate (item, entity) links: summed for GCN/LightGCN, counted
    twice by GraphSage / GAT."""
    n_users, n_items, n_props = 120, 90, 60
    adj = random_bipartite(n_users, n_items, 2500, seed=3, n_props=n_props, n_links=400, dup_links=60)
    assert adj.tocsr().nnz < adj.nnz
    model = _build(name, adj, GRIDS[1])
    u = np.arange(100) % n_users
    i = np.arange(100) % n_items + n_users
    model((u, i))
    _randomise(model, 2)
    w = export_weights(model)
    kind = KINDS[name]
    emb = ol.propagate(kind, w["embeddings"], _oracle_graph(kind, adj), w["layers"])
    assert_close(model.gnn(None).cpu().numpy(), emb, what=name + " uip embeddings")


@pytest.mark.parametrize("name", ["BasicGCN", "BasicGAT"])
def test_full_movielens_shape(name):
    n_users, n_items = 6040, 3706
    adj = random_bipartite(n_users, n_items, 572000, seed=42)
    model = _build(name, adj, GRIDS[1])
    rng = np.random.RandomState(2)
    u = rng.randint(0, n_users, size=2048)
    i = rng.randint(0, n_items, size=2048) + n_users
    model((u, i))
    _randomise(model, 3)
    got = model((u, i)).cpu().numpy()
    w = export_weights(model)
    kind = KINDS[name]
    emb = ol.propagate(kind, w["embeddings"], _oracle_graph(kind, adj), w["layers"])
    assert_close(got, ol.basic_rs(emb, u, i, w["unet"], w["inet"], w["clf"]), what=name + " ML-1M scores")


def test_known_parameter_counts():
    """doc.pdf Table 17, N = 9,228 nodes: pins constructor wiring and weight shapes."""
    adj = random_bipartite(6036, 3192, 50000, seed=1)
    u, i = np.array([0, 1]), np.array([6036, 6037])
    want = {"BasicGCN": 168033, "BasicGraphSage": 168545, "BasicLightGCN": 164417}
    for name, count in want.items():
        model = _build(name, adj, GRIDS[1])
        model((u, i))
        assert model.count_params() == count, name
        assert len(model.non_trainable_weights) == 0
    model = _build("BasicGCN", adj, GRIDS[0])
    model((u, i))
    assert model.count_params() == 81121
    assert model.build_weights().count_params() == 81121


def test_constructor_contract():
    from deep_cbrs_amar_renaissance_b200.models import basic, gnn
    adj = random_bipartite(20, 10, 80, seed=1)
    with pytest.raises(NotImplementedError):
        basic.BasicGCN(adj, cache_neighbours=True)
    with pytest.raises(ValueError):
        basic.BasicGCN(adj, final_node="nope")
    d = basic.BasicDGCF(adj, final_node="concatenation", embedding_dim=8, n_layers=2)
    assert d.gnn.gnn_layers.final_node == "mean"  # gnn.py:405 overrides it
    m = basic.BasicLightGCN(adj, final_node="concatenation", embedding_dim=8, n_layers=2)
    assert m.gnn.gnn_layers.final_node == "mean"  # gnn.py:378 overrides it
    assert len(gnn.GCN(adj, n_hiddens=(8, 8, 8)).gnn_layers) == 3


@pytest.mark.parametrize("final_node", ["sum", "mean", "last", "w-sum"])
def test_other_reductions(final_node):
    adj = random_bipartite(80, 60, 900, seed=4)
    model = _build("BasicGCN", adj, (16, [16, 16], [24], [16]))
    model.gnn.gnn_layers.final_node = final_node
    from deep_cbrs_amar_renaissance_b200.layers import ReductionLayer
    model.gnn.gnn_layers.reduce = ReductionLayer(final_node)
    got = model.gnn(None).cpu().numpy()
    w = export_weights(model)
    want = ol.propagate("gcn", w["embeddings"], og.gcn_filter(adj), w["layers"], final_node=final_node)
    assert_close(got, want, what=final_node)


@pytest.mark.parametrize("name", ["HybridBertGCN", "HybridBertGraphSage", "HybridBertGAT", "HybridBertLightGCN"])
@pytest.mark.parametrize("feature_based", [True, False])
def test_hybrid_models_forward_matches_oracle(name, feature_based):
    n_users, n_items, dim = 150, 100, 768
    adj = random_bipartite(n_users, n_items, 3000, seed=8)
    bert = (np.random.RandomState(5).standard_normal((n_users + n_items, dim)) * 0.5).astype(np.float32)
    grid = (16, [16, 16], [[48, 48], [256, 64], [64, 64]], [64, 64])  # econfigs/hybrid-gnn.yaml grid2
    model = _build(name, adj, grid, module="hybrid", feature_based=feature_based, fusion_method="concatenate",
                   residual=False)
    rng = np.random.RandomState(6)
    u = rng.randint(0, n_users, size=700)
    i = rng.randint(0, n_items, size=700) + n_users
    model((u, i, bert[u], bert[i]))
    _randomise(model, 4)
    got = model((u, i, bert[u], bert[i])).cpu().numpy()  # the reference's call form: host-gathered rows
    w = export_weights(model)
    kind = KINDS[name.replace("HybridBert", "Basic")]
    emb = ol.propagate(kind, w["embeddings"], _oracle_graph(kind, adj), w["layers"])
    want = ol.hybrid_cbrs(emb, u, i, bert[u], bert[i], w, feature_based=feature_based)
    assert_close(got, want, rtol=2e-5, what=name + " scores")  # 768-term fp32 dot products
    model.set_content_table(bert)  # B200 form: ids only, rows gathered in-kernel
    assert_close(model((u, i)).cpu().numpy(), got, rtol=1e-6, what="device-resident content table")


def test_rgcn_extension():
    from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph
    from deep_cbrs_amar_renaissance_b200.models import basic
    n_users, n_items, n_props = 100, 80, 50
    adj = random_bipartite(n_users, n_items, 2000, seed=5, n_props=n_props, n_links=300, dup_links=30)
    dev = torch.device("cuda", 0)
    row, col = torch.from_numpy(adj.row).to(dev), torch.from_numpy(adj.col).to(dev)
    val = torch.from_numpy(adj.data).to(dev)
    n = adj.shape[0]
    # one relation == GCN, bit for bit
    g1 = DeviceGraph(row, col, val, n, rel=torch.zeros_like(row), n_rel=1)
    m1 = _build("BasicRGCN", g1, GRIDS[1])
    m0 = _build("BasicGCN", adj, GRIDS[1])
    u, i = np.arange(64) % n_users, np.arange(64) % n_items + n_users
    m1((u, i)), m0((u, i))
    m1.set_weights(m0.get_weights())
    assert torch.equal(m1.gnn(None), m0.gnn(None))
    # two relations by node range: r0 = user<->item, r1 = item<->property
    rel_np = ((adj.row >= n_users + n_items) | (adj.col >= n_users + n_items)).astype(np.int32)
    g2 = DeviceGraph(row, col, val, n, rel=torch.from_numpy(rel_np).to(dev), n_rel=2)
    m2 = _build("BasicRGCN", g2, GRIDS[1])
    m2((u, i))
    _randomise(m2, 7)
    named = dict((k, v.cpu().numpy()) for k, v in m2.named_weights())
    a_hat = og.gcn_filter(adj).tocoo()
    x = named["gnn/gnn_layers/embeddings"]
    hs = [x]
    from scipy import sparse
    for l in range(2):
        pre = "gnn/gnn_layers/seq_layers.%d/" % l
        rels = []
        for r in range(2):
            is_r1 = (a_hat.row >= n_users + n_items) | (a_hat.col >= n_users + n_items)
            pick = is_r1 if r == 1 else ~is_r1  # self loops (row == col) ride on relation 0 unless a property node
            # self loops are tagged self_rel = 0 by the build
            diag = a_hat.row == a_hat.col
            pick = np.where(diag, r == 0, pick)
            rels.append(sparse.csr_matrix((a_hat.data[pick], (a_hat.row[pick], a_hat.col[pick])), shape=a_hat.shape))
        x = ol.rgcn_conv(x, rels, [named[pre + "kernel_0"], named[pre + "kernel_1"]], named[pre + "bias"])
        hs.append(x)
    assert_close(m2.gnn(None).cpu().numpy(), np.concatenate(hs, 1), what="rgcn")


def test_catalog_top_k_matches_oracle():
    from deep_cbrs_amar_renaissance_b200.scoring import catalog_scores
    n_users, n_items, k = 400, 700, 10
    adj = random_bipartite(n_users, n_items, 9000, seed=12)
    model = _build("BasicGCN", adj, GRIDS[1])
    model((np.array([0]), np.array([n_users])))
    _randomise(model, 5)
    model.cache_propagation = True
    ids, vals = model.recommend_top_k(n_users, n_items, k)                 # fused kernel
    ids_g, vals_g = model.recommend_top_k(n_users, n_items, k, fused=False)  # generic pair pipeline
    w = export_weights(model)
    emb = ol.propagate("gcn", w["embeddings"], og.gcn_filter(adj), w["layers"])
    uu = np.repeat(np.arange(n_users), n_items)
    ii = np.tile(np.arange(n_items), n_users) + n_users
    oracle_scores = ol.basic_rs(emb, uu, ii, w["unet"], w["inet"], w["clf"]).reshape(n_users, n_items)
    clear = assert_topk_equivalent(ids.cpu().numpy(), vals.cpu().numpy(), oracle_scores, k)
    assert clear > 0.95
    dense_scores = catalog_scores(model, model.propagate(), n_users, n_items).cpu().numpy()
    assert_close(dense_scores, oracle_scores, what="catalog scores")
    # the generic path's top-k is bit-exact with a stable sort of its own score matrix
    want_ids, want_vals = ol.top_k_catalog(dense_scores, k)
    assert np.array_equal(ids_g.cpu().numpy(), want_ids) and np.array_equal(vals_g.cpu().numpy(), want_vals)
    assert_topk_equivalent(ids_g.cpu().numpy(), vals_g.cpu().numpy(), oracle_scores, k)
    # a user subset gives the same rows
    sub = torch.tensor([5, 17, 399], device="cuda")
    ids2, _ = model.recommend_top_k(n_users, n_items, k, users=sub)
    assert torch.equal(ids2, ids[sub])


@pytest.mark.parametrize("feature_based", [True, False])
def test_hybrid_catalog_top_k_matches_oracle(feature_based, monkeypatch):
    """Full-catalog top-k of the hybrid scorer (src/models/hybrid.py:72-89 applied to every (user, item)): the
    entity-based form reaches the fused kernel through its hoisted dense3a / dense3b, the feature-based form (what every
    grid of econfigs/hybrid-gnn.yaml uses) the blocked pair pipeline; both against the oracle's scores, and the fused
    result against the generic one."""
    from deep_cbrs_amar_renaissance_b200 import scoring
    n_users, n_items, dim, k = 90, 260, 768, 10
    adj = random_bipartite(n_users, n_items, 4000, seed=18)
    bert = (np.random.RandomState(5).standard_normal((n_users + n_items, dim)) * 0.5).astype(np.float32)
    grid = (16, [16, 16], [[48, 48], [256, 64], [64, 64]], [64, 64])
    model = _build("HybridBertGCN", adj, grid, module="hybrid", feature_based=feature_based, fusion_method="concatenate",
                   residual=False)
    model.set_content_table(bert)
    model((np.array([0]), np.array([n_users])))
    _randomise(model, 9)
    model.cache_propagation = True
    w = export_weights(model)
    emb = ol.propagate("gcn", w["embeddings"], og.gcn_filter(adj), w["layers"])
    uu = np.repeat(np.arange(n_users), n_items)
    ii = np.tile(np.arange(n_items), n_users) + n_users
    oracle_scores = ol.hybrid_cbrs(emb, uu, ii, bert[uu], bert[ii], w, feature_based=feature_based).reshape(n_users, n_items)
    monkeypatch.setattr(scoring._blocks, "__defaults__", (None, 100))   # 100 pairs per block < n_items: item blocks too
    ids_g, vals_g = model.recommend_top_k(n_users, n_items, k, fused=False)
    dense_scores = scoring.catalog_scores(model, model.propagate(), n_users, n_items).cpu().numpy()
    assert_close(dense_scores, oracle_scores, rtol=2e-5, what="hybrid catalog scores")
    want_ids, want_vals = ol.top_k_catalog(dense_scores, k)   # bit-exact with a stable sort of its own scores
    assert np.array_equal(ids_g.cpu().numpy(), want_ids) and np.array_equal(vals_g.cpu().numpy(), want_vals)
    # hybrid scores carry 2e-5 (768-term fp32 dot products): ids are compared where the oracle's gaps exceed that
    assert assert_topk_equivalent(ids_g.cpu().numpy(), vals_g.cpu().numpy(), oracle_scores, k, tol=2e-5) > 0.5
    ids_f, vals_f = model.recommend_top_k(n_users, n_items, k)   # fused where the scorer allows it
    assert assert_topk_equivalent(ids_f.cpu().numpy(), vals_f.cpu().numpy(), oracle_scores, k, tol=2e-5) > 0.5
    if not feature_based:
        assert scoring._hoisted_sources(model, model.propagate(), torch.arange(n_users, device="cuda"),
                                        torch.arange(n_users, n_users + n_items, device="cuda")) is not None


def test_item_blocked_catalog_top_k_keeps_the_tie_rule(monkeypatch):
    """catalogs wider than the pair budget are cut along the item axis and merged: exact ties must still go to the
    lower item id across block borders, and the result must equal the unblocked one"""
    from deep_cbrs_amar_renaissance_b200 import scoring
    n_users, n_items = 30, 50
    adj = random_bipartite(n_users, n_items, 300, seed=2)
    model = _build("BasicGCN", adj, GRIDS[0])
    model((np.array([0]), np.array([n_users])))
    saved = model.get_weights()
    model.cache_propagation = True
    full_ids, full_vals = model.recommend_top_k(n_users, n_items, 7, fused=False)
    monkeypatch.setattr(scoring._blocks, "__defaults__", (None, 16))   # 16 pairs per block: 4 item blocks per user
    ids, vals = model.recommend_top_k(n_users, n_items, 7, fused=False)
    assert torch.equal(ids, full_ids) and torch.equal(vals, full_vals)
    model.set_weights([np.zeros_like(w) for w in saved])  # every score is sigmoid(0) = 0.5
    model.invalidate()
    ids, vals = model.recommend_top_k(n_users, n_items, 7, fused=False)
    assert (ids.cpu().numpy() == np.arange(7)[None, :]).all() and (vals == 0.5).all()


def test_exact_ties_go_to_the_lower_item_index():
    n_users, n_items = 30, 50
    adj = random_bipartite(n_users, n_items, 300, seed=2)
    model = _build("BasicGCN", adj, GRIDS[0])
    model((np.array([0]), np.array([n_users])))
    model.set_weights([np.zeros_like(w) for w in model.get_weights()])  # every score is sigmoid(0) = 0.5
    ids, vals = model.recommend_top_k(n_users, n_items, 7)
    assert (ids.cpu().numpy() == np.arange(7)[None, :]).all() and (vals == 0.5).all()


def test_predict_evaluate_and_pair_top_k(tmp_path):
    """The reference's evaluate(): predict over the test Sequence, then per-user top-5/10 among the
    test pairs (experiment.py:197-207)."""
    from deep_cbrs_amar_renaissance_b200.data import loaders, synthetic
    from deep_cbrs_amar_renaissance_b200.utilities.metrics import top_k_predictions
    paths = synthetic.write_dataset(str(tmp_path), 200, 150, 6000, seed=9)
    train, test = loaders.load_user_item_graph(**paths, train_batch_size=256, test_batch_size=512)
    model = _build("BasicGraphSage", train.adj_matrix, GRIDS[1])
    model.compile(loss="binary_crossentropy", optimizer=None, metrics=["accuracy"])
    model(train[0][0])
    lines = []
    model.summary(print_fn=lines.append, expand_nested=True)
    assert any("Total params" in s for s in lines)
    preds = model.predict(test)
    assert preds.shape == (len(test.ratings), 1)
    loss, acc = model.evaluate(test)
    assert 0 < loss < 5 and 0 <= acc <= 1
    w = export_weights(model)
    ptr, idx, _ = og.reorder_raw(train.adj_matrix)
    emb = ol.propagate("sage", w["embeddings"], (ptr, idx), w["layers"])
    want = ol.basic_rs(emb, test.ratings[:, 0], test.ratings[:, 1], w["unet"], w["inet"], w["clf"])
    assert_close(preds, want, what="predict")
    ratings_pred = np.concatenate([test.ratings[:, [0, 1]], preds], axis=1)
    for k in (5, 10):
        df = top_k_predictions(ratings_pred, train.users, train.items, k=k)
        uu, ii, ss, rows = ol.top_k_pairs(ratings_pred[:, 0].astype(np.int64), ratings_pred[:, 1],
                                          ratings_pred[:, 2].astype(np.float32), k)
        assert df['users'].tolist() == train.users[uu].tolist()
        assert df['items'].tolist() == train.items[ii.astype(np.int64) - len(train.users)].tolist()
    # the reference's Experimenter.train call (experiment.py:183-188) on the golden fixture's Sequence
    model.compile(loss="binary_crossentropy", optimizer={"learning_rate": 1e-3, "beta_1": 0.9}, metrics=["accuracy"])
    hist = model.fit(train, epochs=2, workers=1, callbacks=[])
    assert len(hist.history["loss"]) == 2 and np.isfinite(hist.history["loss"]).all()


def test_large_graph_properties():
    """Sizes the oracle cannot replay quickly: size-independent properties on a 2e7-edge graph, D=128."""
    from deep_cbrs_amar_renaissance_b200 import _lib as L
    from deep_cbrs_amar_renaissance_b200 import ops
    from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph
    dev = torch.device("cuda", 0)
    n_users, n_items, n_edges = 200000, 20000, 10_000_000
    row, col = ops.synth_bipartite(n_users, n_items, n_edges, 42, dev)
    n = n_users + n_items
    g = DeviceGraph(row, col, None, n)
    a = g.norm
    rowptr = a.rowptr
    assert int(rowptr[-1]) == a.nnz and bool((rowptr[1:] >= rowptr[:-1]).all())
    # sortedness + no duplicates inside rows: keys strictly increase
    rows_of = torch.repeat_interleave(torch.arange(n, device=dev), rowptr[1:] - rowptr[:-1])
    keys = rows_of * n + a.colidx.long()
    assert bool((keys[1:] > keys[:-1]).all())
    # checksum: sum of values per (row) of the raw view equals the degree; norm view is symmetric
    deg = (g.raw.rowptr[1:] - g.raw.rowptr[:-1])
    assert int(deg.sum()) == 2 * n_edges and a.chunks["n_heavy"] > 0
    ones = torch.ones(n, 128, device=dev)
    out = torch.empty(n, 128, device=dev)
    ops.spmm(g.raw, ones, out, agg=L.AGG_SUM)
    assert torch.equal(out[:, 0], deg.float()) and torch.equal(out[:, 0], out[:, 127])
    # linearity: A(x + 2y) == Ax + 2Ay within fp32 rounding
    x = torch.randn(n, 128, device=dev)
    y = torch.randn(n, 128, device=dev)
    ax, ay, axy = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    ops.spmm(a, x, ax), ops.spmm(a, y, ay), ops.spmm(a, x + 2 * y, axy)
    err = (axy - (ax + 2 * ay)).abs().max().item()
    assert err <= 1e-5 * axy.abs().max().item()
    # idempotence of the build and determinism of the kernel
    again = torch.empty_like(ax)
    ops.spmm(a, x, again)
    assert torch.equal(again, ax)
    # symmetry: x^T (A y) == (A x)^T y in float64
    lhs = (x.double() * ay.double()).sum().item()
    rhs = (ax.double() * y.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)  # fp32 outputs, float64 reduction


@pytest.mark.parametrize("case", ["ui", "uip-dups"])
def test_dgcf_operator_matches_the_reference_recipe(case):
    """DGCFConv.preprocess on device (SpGEMM products + the graph-build pipeline + threshold search) against the
    scipy recipe of dgcf_conv.py:38-80: same epsilon chosen, same structure, values within 1e-5."""
    from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph
    if case == "ui":
        adj = random_bipartite(150, 100, 1800, seed=5)
    else:
        adj = random_bipartite(120, 90, 1500, seed=6, n_props=40, n_links=300, dup_links=60)
    want, info = ol.dgcf_preprocess(adj)
    g = DeviceGraph.from_scipy(adj)
    got = g.dgcf.to_scipy()
    assert g.dgcf_info["epsilon"] == info["epsilon"] and g.dgcf_info["edges"] == info["edges"]
    assert g.dgcf_info["cross_edges"] == info["cross_edges"], (g.dgcf_info, info)
    got.sort_indices()
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert_close(got.data, want.data, what="dgcf operator values")
    assert abs(got - got.T).max() < 1e-6  # symmetric: the training backward relies on it
