"""GPU parity of the training step (scope row (f)-1): the explicit backward kernels, the loss,
the l2 penalty and the Adam update against the torch-CPU autograd oracle (oracle/train.py, float64).

Tolerance: gradients 2e-4 relative to the largest gradient entry of the same tensor (fp32 kernels vs a
float64 oracle; sums over up to 1e4 rows), loss 1e-5 (2e-5 with the l2 penalty), each Adam
update 2e-6 absolute against the written-out Keras formula."""
import numpy as np
import pytest
import torch

from oracle import graph as og
from oracle import train as ot
from tests.helpers import assert_close, export_weights, random_bipartite
from tests.test_gpu_models import KINDS, _build, _oracle_graph, _randomise

pytestmark = pytest.mark.gpu

TRAINABLE = ["BasicGCN", "BasicGraphSage", "BasicLightGCN", "BasicGAT", "BasicDGCF"]
GRAD_RTOL = 2e-4
GRAD_FLOOR = 1e-10  # a gradient that is analytically ~0 (e.g. GAT's attn_kernel_self when all scores share a sign)


def assert_grad_close(got, want, what):
    want = np.asarray(want, np.float64)
    tol = max(GRAD_RTOL * np.abs(want).max(), GRAD_FLOOR) if want.size else GRAD_FLOOR
    assert_close(got, want, atol=tol, what=what)


@pytest.fixture(scope="module", autouse=True)
def _device():
    from deep_cbrs_amar_renaissance_b200 import ops
    ops.check_device()
    torch.cuda.set_device(0)


def _named_grads(model, tape):
    """oracle leaf name -> product gradient (numpy)"""
    out = {}
    for name, w in model.named_weights():
        g = tape.wgrads.get(id(w))
        if g is None:
            continue
        g = g.detach().cpu().numpy()
        if name == "gnn/gnn_layers/embeddings":
            out["embeddings"] = g
        elif name.startswith("gnn/gnn_layers/seq_layers."):
            k, leaf = name[len("gnn/gnn_layers/seq_layers."):].split("/", 1)
            out["layers.%s.%s" % (k, leaf)] = g.reshape(g.shape[0], -1) if leaf == "kernel" else g.reshape(-1) if leaf.startswith("attn") else g
        elif name.startswith("rs/fuse"):
            out[name[3:].replace("/", ".")] = g
        elif name.startswith("rs/"):
            stack, layer, leaf = name[3:].split("/")
            out["%s.%s.%s" % (stack, layer.split(".")[1], leaf)] = g
    return out


def _batch(n_users, n_items, b, seed):
    rng = np.random.RandomState(seed)
    u = rng.randint(0, n_users, size=b)
    u[: b // 8] = u[0]  # repeated ids: the lookup backward must add duplicates
    i = rng.randint(0, n_items, size=b) + n_users
    y = rng.randint(0, 2, size=b)
    return u, i, y


@pytest.mark.parametrize("name", TRAINABLE)
@pytest.mark.parametrize("final_node", ["concatenation", "mean"])
def test_gradients_match_autograd_oracle(name, final_node):
    from deep_cbrs_amar_renaissance_b200 import training
    n_users, n_items = 300, 200
    adj = random_bipartite(n_users, n_items, 6000, seed=7)
    model = _build(name, adj, (16, [16, 16], [48, 48], [64, 64]))
    if KINDS[name] in ("lightgcn", "dgcf"):
        if final_node != "mean":
            pytest.skip("LightGCN and DGCF force final_node='mean' (gnn.py:378,405)")
    else:
        from deep_cbrs_amar_renaissance_b200.layers import ReductionLayer
        model.gnn.gnn_layers.final_node = final_node
        model.gnn.gnn_layers.reduce = ReductionLayer(final_node)
    u, i, y = _batch(n_users, n_items, 1024, 3)
    model((u, i))
    _randomise(model, seed=5)
    w = export_weights(model)
    kind = KINDS[name]
    tape, loss, correct, probs = training.forward_backward(model, (u, i), y)
    torch.cuda.synchronize()
    fn = "mean" if kind in ("lightgcn", "dgcf") else final_node
    want, want_loss, want_p = ot.gradients(kind, w, _oracle_graph(kind, adj), (u, i), y, final_node=fn)
    assert_close(probs.cpu().numpy().reshape(-1), want_p, rtol=2e-5, what=name + " probabilities")
    assert abs(float(loss.item()) - want_loss) <= 1e-5 * max(1.0, abs(want_loss))
    assert int(correct.item()) == int(((want_p > 0.5) == (y > 0.5)).sum())
    got = _named_grads(model, tape)
    assert set(got) == set(want), (sorted(got), sorted(want))
    for k in sorted(want):
        assert_grad_close(got[k], want[k], "%s grad %s" % (name, k))


def test_hybrid_gradients_match_autograd_oracle():
    from deep_cbrs_amar_renaissance_b200 import training
    n_users, n_items = 200, 150
    adj = random_bipartite(n_users, n_items, 4000, seed=9)
    model = _build("HybridBertGCN", adj, (16, [16, 16], [[48, 48], [96, 32], [64, 64]], [64, 64]), module="hybrid",
                   feature_based=True)
    rng = np.random.RandomState(2)
    u, i, y = _batch(n_users, n_items, 512, 4)
    ub = (rng.standard_normal((512, 96)) * 0.5).astype(np.float32)
    ib = (rng.standard_normal((512, 96)) * 0.5).astype(np.float32)
    model((u, i, ub, ib))
    _randomise(model, seed=6)
    w = export_weights(model)
    tape, loss, correct, probs = training.forward_backward(model, (u, i, ub, ib), y)
    want, want_loss, want_p = ot.gradients("gcn", w, og.gcn_filter(adj), (u, i, ub, ib), y, hybrid=True, feature_based=True)
    assert_close(probs.cpu().numpy().reshape(-1), want_p, rtol=2e-5, what="hybrid probabilities")
    got = _named_grads(model, tape)
    assert set(got) == set(want)
    for k in sorted(want):
        assert_grad_close(got[k], want[k], "hybrid grad %s" % k)


@pytest.mark.parametrize("name", TRAINABLE)
def test_three_adam_steps_match_oracle(name):
    """3 optimiser steps on 3 batches.  At every step the oracle restarts from the product's current
    weights (a float64 trajectory of its own drifts away through relu kinks and Adam's division by
    sqrt(v), which says nothing about the kernels), and three things are checked: the loss including the
    l2 penalty, every gradient at the evolved weights, and the Adam update itself against the written-out
    Keras formula applied to the product's gradient + 2*l2*w."""
    from deep_cbrs_amar_renaissance_b200 import training
    n_users, n_items = 300, 200
    adj = random_bipartite(n_users, n_items, 6000, seed=7)
    model = _build(name, adj, (8, [8, 8], [24, 24], [48, 48]))  # l2_regularizer=1e-4 in _build
    batches = [_batch(n_users, n_items, 512, 10 + s) for s in range(3)]
    model((batches[0][0], batches[0][1]))
    _randomise(model, seed=8)
    kind = KINDS[name]
    graph = _oracle_graph(kind, adj)
    fn = "mean" if kind in ("lightgcn", "dgcf") else "concatenation"
    adam = training.Adam(learning_rate=1e-2)
    l2 = training.l2_coefficients(model)
    # embeddings + per layer: (kernel, bias) for GCN/GAT/GraphSage, the gate for DGCF, nothing for LightGCN
    assert len(l2) == {"lightgcn": 1, "dgcf": 3}.get(kind, 5) and set(l2.values()) == {1e-4}
    state = {}
    for t, (u, i, y) in enumerate(batches, start=1):
        w = export_weights(model)
        tape, loss, _, _ = training.forward_backward(model, (u, i), y)
        ws = [x for x in model.trainable_weights if id(x) in tape.wgrads]
        for x in ws:
            if l2.get(id(x)):
                from deep_cbrs_amar_renaissance_b200 import ops
                ops.sum_squares(x, l2[id(x)], loss, accumulate=True)
        want, want_loss, _ = ot.gradients(kind, w, graph, (u, i), y, l2=1e-4, final_node=fn)
        assert abs(float(loss.item()) - want_loss) <= 2e-5 * max(1.0, abs(want_loss)), (t, float(loss.item()), want_loss)
        before = {n: x.detach().cpu().numpy().astype(np.float64) for n, x in model.named_weights()}
        prod_g = {n: tape.wgrads[id(x)].detach().cpu().numpy().astype(np.float64) for n, x in model.named_weights()}
        # gradient parity at the evolved weights (the oracle's gradient includes 2*l2*w; the product adds it in Adam)
        got = _named_grads(model, tape)
        names = {n: k for n, k in zip([n for n, _ in model.named_weights()], _oracle_names(model))}
        for n, x in model.named_weights():
            full = prod_g[n] + 2 * l2.get(id(x), 0.0) * before[n]
            assert_grad_close(full, want[names[n]].reshape(full.shape), "%s step %d grad %s" % (name, t, n))
        adam.apply(ws, [tape.wgrads[id(x)] for x in ws], [l2.get(id(x), 0.0) for x in ws])
        torch.cuda.synchronize()
        for n, x in model.named_weights():
            g = prod_g[n] + 2 * l2.get(id(x), 0.0) * before[n]
            m, v = state.get(n, (np.zeros_like(g), np.zeros_like(g)))
            new, m, v = ot.adam_update(before[n], g, m, v, t, lr=1e-2)
            state[n] = (m, v)
            assert_close(x.detach().cpu().numpy(), new, atol=2e-6, what="%s step %d adam %s" % (name, t, n))
    assert adam.iterations == 3


def _oracle_names(model):
    out = []
    for name, _ in model.named_weights():
        if name == "gnn/gnn_layers/embeddings":
            out.append("embeddings")
        elif name.startswith("gnn/gnn_layers/seq_layers."):
            k, leaf = name[len("gnn/gnn_layers/seq_layers."):].split("/", 1)
            out.append("layers.%s.%s" % (k, leaf))
        else:
            stack, layer, leaf = name[3:].split("/")
            out.append("%s.%s.%s" % (stack, layer.split(".")[1], leaf))
    return out


def test_train_on_batch_and_compile_surface():
    """model.compile(...) + train_on_batch: the optimiser config forms the reference passes"""
    from deep_cbrs_amar_renaissance_b200 import training
    adj = random_bipartite(60, 40, 600, seed=1)
    model = _build("BasicGCN", adj, (8, [8, 8], [24, 24], [48, 48]))
    u, i, y = _batch(60, 40, 128, 0)
    model.compile(loss="binary_crossentropy", optimizer=training.Adam(learning_rate=1e-3, beta_1=0.9), metrics=["accuracy"])
    l0, c0 = model.train_on_batch((u, i), y)
    for _ in range(20):
        l1, _ = model.train_on_batch((u, i), y)
    assert float(l1.item()) < float(l0.item())
    assert 0 <= int(c0.item()) <= 128
    with pytest.raises(NotImplementedError):
        model.compile(loss="mse", optimizer="adam")


def _flat_views(w):
    out = {"embeddings": w["embeddings"]}
    for li, lw in enumerate(w["layers"]):
        for leaf in ("kernel", "bias"):
            if leaf in lw:
                out["layers.%d.%s" % (li, leaf)] = lw[leaf]
    for stack in ("unet", "inet", "clf", "dense1a", "dense1b", "dense2a", "dense2b", "dense3a", "dense3b"):
        for k, (kern, bias) in enumerate(w.get(stack, [])):
            out["%s.%d.kernel" % (stack, k)] = kern
            out["%s.%d.bias" % (stack, k)] = bias
    return out


def test_fit_lowers_the_loss_and_fires_callbacks():
    """Keras-like fit on the reference's Sequence: loss goes down on a learnable synthetic signal, the
    callback hooks of src/utilities/keras.py fire, the dataset reshuffles at epoch end."""
    from deep_cbrs_amar_renaissance_b200.data.datasets import UserItemGraph
    n_users, n_items = 120, 80
    rng = np.random.RandomState(0)
    uu = rng.randint(0, n_users, size=4000)
    ii = rng.randint(0, n_items, size=4000)
    yy = ((uu % 2) == (ii % 2)).astype(np.int64)  # parity rule: learnable from ids
    ratings = np.stack([uu, ii + n_users, yy], axis=1)
    ratings = np.unique(ratings, axis=0)
    users, items = np.arange(n_users), np.arange(n_items)
    from deep_cbrs_amar_renaissance_b200.data.preprocess import build_adjacency_matrix
    ds = UserItemGraph(ratings, users, items, build_adjacency_matrix(ratings, users, items), batch_size=256, shuffle=True)
    model = _build("BasicGCN", ds.adj_matrix, (16, [16, 16], [48, 48], [64, 64]))
    model.compile(loss="binary_crossentropy", optimizer={"learning_rate": 1e-2}, metrics=["accuracy"])
    events = []

    class Spy:
        def on_train_begin(self, logs=None): events.append("train_begin")
        def on_epoch_end(self, epoch, logs=None): events.append(("epoch_end", epoch, dict(logs)))
        def on_train_end(self, logs=None): events.append("train_end")

    hist = model.fit(ds, epochs=10, workers=1, callbacks=[Spy()])
    losses = hist.history["loss"]
    assert len(losses) == 10 and losses[-1] < losses[0] - 0.02, losses
    assert hist.history["accuracy"][-1] > hist.history["accuracy"][0]
    assert events[0] == "train_begin" and events[-1] == "train_end" and sum(isinstance(e, tuple) for e in events) == 10
    ev = model.evaluate(ds)
    assert ev[0] < losses[0]


def test_gat_on_uip_graph_with_duplicates_and_self_loops():
    """duplicate (item, entity) links are separate edges for GAT, existing self loops are replaced: the
    backward must treat both exactly like the forward"""
    from deep_cbrs_amar_renaissance_b200 import training
    from scipy import sparse
    adj = random_bipartite(120, 90, 2500, seed=3, n_props=60, n_links=400, dup_links=80)
    n = adj.shape[0]
    loops = np.arange(0, n, 7, dtype=np.int32)  # some explicit self loops in the input
    adj = sparse.coo_matrix((np.concatenate([adj.data, np.ones(len(loops), np.float32)]),
                             (np.concatenate([adj.row, loops]), np.concatenate([adj.col, loops]))), shape=adj.shape)
    model = _build("BasicGAT", adj, (8, [8, 8], [24, 24], [48, 48]))
    u, i, y = _batch(120, 90, 256, 1)
    model((u, i))
    _randomise(model, seed=2)
    w = export_weights(model)
    tape, loss, _, probs = training.forward_backward(model, (u, i), y)
    want, want_loss, want_p = ot.gradients("gat", w, _oracle_graph("gat", adj), (u, i), y)
    assert_close(probs.cpu().numpy().reshape(-1), want_p, rtol=2e-5, what="gat-uip probabilities")
    got = _named_grads(model, tape)
    assert set(got) == set(want)
    for k in sorted(want):
        assert_grad_close(got[k], want[k], "gat-uip grad %s" % k)


@pytest.mark.parametrize("name", ["BasicGCN", "BasicGraphSage"])
def test_cuda_graph_replay_equals_eager_steps(name):
    """5 optimiser steps replayed from the captured graph leave exactly the weights 5 eager steps leave"""
    n_users, n_items = 300, 200
    adj = random_bipartite(n_users, n_items, 6000, seed=7)
    batches = [_batch(n_users, n_items, 256, 20 + s) for s in range(5)]
    finals = []
    for graphed in (False, True):
        model = _build(name, adj, (8, [8, 8], [24, 24], [48, 48]))
        model.compile(loss="binary_crossentropy", optimizer={"learning_rate": 1e-2})
        model.build_weights()
        step = model.make_graphed_train_step(256) if graphed else None
        losses = []
        for u, i, y in batches:
            loss, _ = step((u, i), y) if graphed else model.train_on_batch((u, i), y)
            losses.append(float(loss.item()))
        finals.append((losses, [w.copy() for w in model.get_weights()]))
    assert finals[0][0] == finals[1][0], (finals[0][0], finals[1][0])
    for a, b in zip(finals[0][1], finals[1][1]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("tweak", ["attention", "residual", "attention-projected"])
def test_hybrid_tweaks_forward_and_gradients(tweak):
    """econfigs/hybrid-gnn-tweaks*.yaml: attention fusion of the two branches, residual classifier; forward against
    the numpy oracle, gradients against autograd"""
    from deep_cbrs_amar_renaissance_b200 import training
    from oracle import layers as ol
    n_users, n_items = 200, 150
    adj = random_bipartite(n_users, n_items, 4000, seed=9)
    units = [[48, 48], [96, 32], [64, 64]] if tweak != "attention-projected" else [[48, 48], [96, 32], [64, 64]]
    extra = dict(feature_based=True, fusion_method="attention" if tweak.startswith("attention") else "concatenate",
                 residual=(tweak == "residual"))
    if tweak == "attention-projected":  # entity-based: fuse1a/1b see widths 48 and 32 -> the narrower is projected
        extra["feature_based"] = False
    model = _build("HybridBertGCN", adj, (16, [16, 16], units, [64, 64]), module="hybrid", **extra)
    rng = np.random.RandomState(2)
    u, i, y = _batch(n_users, n_items, 256, 4)
    ub = (rng.standard_normal((256, 96)) * 0.5).astype(np.float32)
    ib = (rng.standard_normal((256, 96)) * 0.5).astype(np.float32)
    model((u, i, ub, ib))
    _randomise(model, seed=6)
    w = export_weights(model)
    emb = ol.propagate("gcn", w["embeddings"], og.gcn_filter(adj), w["layers"])
    want_fwd = ol.hybrid_cbrs(emb, u, i, ub, ib, w, feature_based=extra["feature_based"])
    assert_close(model((u, i, ub, ib)).cpu().numpy(), want_fwd, rtol=2e-5, what=tweak + " forward")
    tape, loss, correct, probs = training.forward_backward(model, (u, i, ub, ib), y)
    want, want_loss, want_p = ot.gradients("gcn", w, og.gcn_filter(adj), (u, i, ub, ib), y, hybrid=True,
                                           feature_based=extra["feature_based"])
    assert_close(probs.cpu().numpy().reshape(-1), want_p, rtol=2e-5, what=tweak + " probabilities")
    got = _named_grads(model, tape)
    assert set(got) == set(want), (sorted(set(got) ^ set(want)))
    for k in sorted(want):
        assert_grad_close(got[k], want[k], "%s grad %s" % (tweak, k))
    # catalog scoring goes through the same tail
    model.set_content_table((rng.standard_normal((n_users + n_items, 96)) * 0.5).astype(np.float32))
    ids, vals = model.recommend_top_k(n_users, n_items, 5, users=torch.arange(8, device="cuda"))
    assert ids.shape == (8, 5) and (vals[:, :-1] >= vals[:, 1:]).all()
