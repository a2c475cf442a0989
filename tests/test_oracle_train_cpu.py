"""The training oracle checked against itself (host only): oracle/train.py's autograd gradients against central finite
differences of its own loss in float64, for every layer family and the hybrid tweaks, on a tiny graph.  This pins the
differentiation (what the product's backward kernels are compared with on the GPU), not the forward semantics - those
are covered by tests/test_oracle_identities.py and the golden fixtures."""
import numpy as np
import pytest

from oracle import graph as og
from oracle import layers as ol
from oracle import train as ot
from tests.helpers import glorot, random_bipartite


def _weights(kind, rng, n, d=6, hybrid=False, tweak=None):
    def stack(din, units):
        out = []
        for u in units:
            out.append((glorot(rng, (din, u)), (rng.standard_normal(u) * 0.1).astype(np.float32)))
            din = u
        return out, din

    layers = []
    for _ in range(2):
        if kind == "gcn":
            layers.append(dict(kernel=glorot(rng, (d, d)), bias=(rng.standard_normal(d) * 0.1).astype(np.float32)))
        elif kind == "sage":
            layers.append(dict(kernel=glorot(rng, (2 * d, d)), bias=(rng.standard_normal(d) * 0.1).astype(np.float32)))
        elif kind == "gat":
            layers.append(dict(kernel=glorot(rng, (d, d)), bias=(rng.standard_normal(d) * 0.1).astype(np.float32),
                               attn_self=glorot(rng, (d, 1)).reshape(-1), attn_neigh=glorot(rng, (d, 1)).reshape(-1)))
        elif kind == "rgcn":
            layers.append(dict(kernel_0=glorot(rng, (d, d)), kernel_1=glorot(rng, (d, d)),
                               bias=(rng.standard_normal(d) * 0.1).astype(np.float32)))
        elif kind == "dgcf":
            layers.append({"locality_adaptive/locality-adaptive-weights": (1 + 0.3 * rng.standard_normal((n, 1))).astype(np.float32)})
        else:
            layers.append({})
    w = dict(embeddings=glorot(rng, (n, d)), layers=layers)
    d_out = d if kind in ("lightgcn", "dgcf") else 3 * d
    if not hybrid:
        w["unet"], du = stack(d_out, [8])
        w["inet"], di = stack(d_out, [8])
        w["clf"], _ = stack(du + di, [8, 1])
        return w
    w["dense1a"], g = stack(d_out, [8])
    w["dense1b"], _ = stack(d_out, [8])
    w["dense2a"], c = stack(10, [6])
    w["dense2b"], _ = stack(10, [6])
    w["dense3a"], o1 = stack(2 * g, [8])
    w["dense3b"], o2 = stack(2 * c, [8])
    if tweak == "attention":
        w["fuse2"] = dict(att_weight=glorot(rng, (8, 8)), proj_weight=None, proj_first=None)
        w["clf"], _ = stack(8, [8, 1])
    elif tweak == "residual":
        w["residual"], _ = stack(o1 + o2, [8, 8])
        w["clf"], _ = stack(8, [1])
    else:
        w["clf"], _ = stack(o1 + o2, [8, 1])
    return w


def _to64(x):
    if isinstance(x, np.ndarray):
        return x.astype(np.float64)
    if isinstance(x, dict):
        return {k: _to64(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to64(v) for v in x)
    return x


def _graph(kind, adj):
    if kind in ("gcn", "lightgcn"):
        return og.gcn_filter(adj)
    if kind == "rgcn":   # two relations by node range: edges touching the last third of the nodes are relation 1
        from tests.helpers import relation_blocks
        return relation_blocks(og.gcn_filter(adj), adj.shape[0] - adj.shape[0] // 3)
    if kind == "dgcf":
        return ol.dgcf_preprocess(adj)[0]
    ptr, idx, _ = og.reorder_raw(adj)
    return (ptr, idx)


CASES = [("gcn", False, None), ("rgcn", False, None), ("sage", False, None), ("gat", False, None), ("lightgcn", False, None), ("dgcf", False, None),
         ("gcn", True, None), ("gcn", True, "attention"), ("gcn", True, "residual")]


@pytest.mark.parametrize("kind,hybrid,tweak", CASES)
def test_autograd_matches_finite_differences(kind, hybrid, tweak):
    rng = np.random.RandomState(3)
    n_users, n_items = 12, 9
    adj = random_bipartite(n_users, n_items, 60, seed=2)
    n = n_users + n_items
    w = _to64(_weights(kind, rng, n, hybrid=hybrid, tweak=tweak))  # float64 storage: the step can be tiny (relu kinks)
    u = rng.randint(0, n_users, 16)
    i = rng.randint(0, n_items, 16) + n_users
    y = rng.randint(0, 2, 16)
    inputs = (u, i) if not hybrid else (u, i, rng.standard_normal((16, 10)).astype(np.float32), rng.standard_normal((16, 10)).astype(np.float32))
    kw = dict(l2=1e-3, hybrid=hybrid, feature_based=True)
    graph = _graph(kind, adj)
    grads, loss, _ = ot.gradients(kind, w, graph, inputs, y, **kw)

    def loss_at():
        return float(ot.forward_loss(kind, w, graph, inputs, y, **kw)[0].detach())

    # probe a few entries of a few tensors (the embeddings always, plus every other leaf kind present)
    probes = [("embeddings", w["embeddings"])]
    lw = w["layers"][0]
    for key, leaf in (("kernel", "layers.0.kernel"), ("kernel_1", "layers.0.kernel_1"), ("bias", "layers.0.bias"),
                      ("attn_self", "layers.0.attn_kernel_self"),
                      ("locality_adaptive/locality-adaptive-weights", "layers.0.locality_adaptive/locality-adaptive-weights")):
        if key in lw:
            probes.append((leaf, lw[key]))
    probes.append(("clf.0.kernel", w["clf"][0][0]))
    if tweak == "attention":
        probes.append(("fuse2.att_weight", w["fuse2"]["att_weight"]))
    if tweak == "residual":
        probes.append(("residual.1.kernel", w["residual"][1][0]))
    eps = 1e-6
    for name, arr in probes:
        flat = arr.reshape(-1)
        for j in rng.choice(flat.size, size=min(4, flat.size), replace=False):
            old = flat[j]
            flat[j] = old + eps
            hi = loss_at()
            step_up = float(flat[j]) - float(old)
            flat[j] = old - eps
            lo = loss_at()
            step_dn = float(old) - float(flat[j])
            flat[j] = old
            fd = (hi - lo) / (step_up + step_dn)
            g = grads[name].reshape(-1)[j]
            assert abs(fd - g) <= 1e-4 * max(abs(g), 1e-3), (kind, tweak, name, int(j), fd, g)


def _kg_parts(kind, rng, two_way, graphs, n_users, n_items, d=6, side_node="mean"):
    """weights of every SequentialGNN of a Two-Step / Two-Way model (float64) + the scorer"""
    ui, ip, up = graphs
    n_props = ip.shape[0] - n_items
    d2 = d if (side_node == "mean" or kind in ("lightgcn", "dgcf")) else 3 * d

    def part(n, width, adj, with_emb=True, n_emb=None):
        w = _weights(kind, rng, n, d=width)
        return dict(embeddings=glorot(rng, (n_emb or n, width)) if with_emb else None, layers=w["layers"], graph=_graph(kind, adj))

    if two_way:
        parts = dict(way_one=part(n_users + n_props, d, up), way_two=part(n_items + n_props, d, ip),
                     step_two=part(n_users + n_items, d2, ui, with_emb=False))
    else:
        parts = dict(step_one=part(n_items + n_props, d, ip), step_two=part(n_users + n_items, d2, ui, n_emb=n_users))
    d_out = d2 if kind in ("lightgcn", "dgcf") else 3 * d2
    scorer = _weights("lightgcn", rng, 4, d=d_out)   # 'lightgcn' -> towers sized for a d_out-wide table
    return _to64({k: {kk: vv for kk, vv in v.items() if kk != "graph"} for k, v in parts.items()}), \
        {k: v["graph"] for k, v in parts.items()}, _to64({k: scorer[k] for k in ("unet", "inet", "clf")})


@pytest.mark.parametrize("kind,two_way,side_node", [("gcn", False, "mean"), ("gcn", False, "concatenation"), ("sage", False, "mean"),
                                                    ("gat", True, "mean"), ("gcn", True, "concatenation"),
                                                    ("lightgcn", True, "mean"), ("dgcf", False, "mean")])
def test_kg_autograd_matches_finite_differences(kind, two_way, side_node):
    """the Two-Step / Two-Way training oracle (forward_loss_kg) against central differences of its own loss"""
    from deep_cbrs_amar_renaissance_b200.data.preprocess import get_user_properties
    from tests.helpers import kg_graphs
    rng = np.random.RandomState(5)
    n_users, n_items, n_props = 12, 9, 7
    ui, ip = kg_graphs(n_users, n_items, n_props, n_pos=40, n_links=14, dup_links=2)
    graphs = (ui, ip, get_user_properties(ui, ip, n_users, n_items))
    parts, gr, w = _kg_parts(kind, rng, two_way, graphs, n_users, n_items, side_node=side_node)
    for k in parts:
        parts[k]["graph"] = gr[k]
    u = rng.randint(0, n_users, 16)
    i = rng.randint(0, n_items, 16) + n_users
    y = rng.randint(0, 2, 16)
    kw = dict(l2=1e-3, side_node=side_node)
    grads, loss, _ = ot.gradients_kg(kind, parts, w, (u, i), y, n_users, n_items, **kw)
    first = "way_one" if two_way else "step_one"
    probes = [(first + ".embeddings", parts[first]["embeddings"])]
    if two_way:
        probes.append(("way_two.embeddings", parts["way_two"]["embeddings"]))
    else:
        probes.append(("step_two.embeddings", parts["step_two"]["embeddings"]))
    for part in (first, "step_two"):
        lw = parts[part]["layers"][0]
        for key, leaf in (("kernel", "kernel"), ("attn_neigh", "attn_kernel_neigh"),
                          ("locality_adaptive/locality-adaptive-weights", "locality_adaptive/locality-adaptive-weights")):
            if key in lw:
                probes.append(("%s.layers.0.%s" % (part, leaf), lw[key]))
    assert set(grads) >= {name for name, _ in probes}
    eps = 1e-6
    for name, arr in probes:
        flat = arr.reshape(-1)
        for j in rng.choice(flat.size, size=min(4, flat.size), replace=False):
            old = flat[j]
            flat[j] = old + eps
            hi = float(ot.forward_loss_kg(kind, parts, w, (u, i), y, n_users, n_items, **kw)[0].detach())
            flat[j] = old - eps
            lo = float(ot.forward_loss_kg(kind, parts, w, (u, i), y, n_users, n_items, **kw)[0].detach())
            flat[j] = old
            fd = (hi - lo) / (2 * eps)
            g = grads[name].reshape(-1)[j]
            assert abs(fd - g) <= 1e-4 * max(abs(g), 1e-3), (kind, two_way, name, int(j), fd, g)
    # rows of the side graphs that the next step never reads (properties) still get gradient through propagation,
    # user rows of the step-two embeddings get theirs directly
    assert np.abs(grads[first + ".embeddings"]).max() > 0


def test_adam_update_is_the_keras_formula():
    w, g = np.array([1.0, -2.0]), np.array([0.5, -0.25])
    m = v = np.zeros(2)
    w1, m1, v1 = ot.adam_update(w, g, m, v, 1, lr=1e-3)
    # first step of Adam moves every weight by lr * g/|g| (up to epsilon)
    assert np.allclose(w - w1, 1e-3 * np.sign(g), atol=1e-8)
    assert np.allclose(m1, 0.1 * g) and np.allclose(v1, 0.001 * g * g)


def test_bf16_round_is_round_to_nearest_even():
    x = np.array([1.0, 1.00390625, 1.01171875, -3.1415927, 0.0, 65504.0], np.float32)
    r = ol.bf16_round(x)
    assert r[0] == 1.0 and r[1] == 1.0 and r[2] == np.float32(1.015625)   # ties go to the even mantissa
    import torch
    assert np.array_equal(r, torch.from_numpy(x).to(torch.bfloat16).float().numpy())


def test_dgcf_operator_properties():
    adj = random_bipartite(40, 30, 300, seed=4)
    m, info = ol.dgcf_preprocess(adj)
    assert info["epsilon"] in ol.DGCF_EPSILONS and abs(m - m.T).max() < 1e-6
    a_hat = og.gcn_filter(adj)
    rest = (m - a_hat - __import__("scipy.sparse", fromlist=["eye"]).eye(m.shape[0], format="csr")).tocsr()
    rest.eliminate_zeros()
    assert rest.nnz == 0 or rest.data.min() > info["epsilon"] - 1e-6  # what remains is the filtered crosshop
