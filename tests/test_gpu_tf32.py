"""GPU parity for the fp32-accurate tensor-core transform (cbrs_dense_tf32x3: three TF32 tcgen05 MMAs per product, X
tiles by TMA).  Oracle: the float64 product rounded once.  Tolerance: 1e-5 of the output scale, the bound the fp32
path promises (north star); the kernel is expected to sit near 1e-6, which the comparison with the FFMA kernel checks
with a tighter, stated bound."""
import numpy as np
import pytest
import torch

from oracle import graph as og
from tests.helpers import assert_close, random_bipartite

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from deep_cbrs_amar_renaissance_b200 import ops
    ops.check_device()
    return torch.device("cuda", 0)


def _case(m, k, n, seed, dev, scale=1.0):
    rng = np.random.RandomState(seed)
    x = (rng.standard_normal((m, k)) * scale).astype(np.float32)
    w = (rng.standard_normal((k, n)) / np.sqrt(k)).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    return x, w, b, torch.from_numpy(x).to(dev), torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev)


@pytest.mark.parametrize("m,k,n", [(1000, 128, 128), (129, 32, 16), (4101, 64, 256), (300, 256, 64), (77, 96, 48),
                                   (1, 128, 128), (128, 128, 128), (148 * 128 * 2 + 3, 128, 128)])
def test_tf32x3_matches_float64_product(dev, m, k, n):
    from deep_cbrs_amar_renaissance_b200 import ops
    x, w, b, xd, wd, bd = _case(m, k, n, m + k + n, dev)
    want = x.astype(np.float64) @ w.astype(np.float64)
    got = ops.dense_tf32x3(xd, wd)
    assert_close(got.cpu().numpy(), want.astype(np.float32), rtol=1e-5, what="x @ w (3xTF32)")
    got = ops.dense_tf32x3(xd, wd, bd, "relu")
    assert_close(got.cpu().numpy(), np.maximum(want + b, 0).astype(np.float32), rtol=1e-5, what="relu(x @ w + b)")
    # against the fp32 FFMA kernel (its own error vs float64 is ~1e-7 of the scale): 3e-6 of the scale
    ffma = ops.dense(xd, wd)
    scale = float(ffma.abs().max())
    assert float((ops.dense_tf32x3(xd, wd) - ffma).abs().max()) <= 3e-6 * scale


def test_tf32x3_wide_dynamic_range_and_exact_cases(dev):
    """entries spanning 2^-20 .. 2^20 in one row; and operands that are exactly representable in tf32 (small integers):
    the split's low halves are zero and the result must be exact"""
    from deep_cbrs_amar_renaissance_b200 import ops
    rng = np.random.RandomState(3)
    x = (rng.standard_normal((513, 128)) * np.exp2(rng.randint(-20, 20, size=(513, 128)))).astype(np.float32)
    w = (rng.standard_normal((128, 128)) / 11).astype(np.float32)
    want = (x.astype(np.float64) @ w.astype(np.float64)).astype(np.float32)
    got = ops.dense_tf32x3(torch.from_numpy(x).to(dev), torch.from_numpy(w).to(dev)).cpu().numpy()
    assert_close(got, want, rtol=1e-5, what="wide dynamic range")
    xi = rng.randint(-8, 9, size=(300, 64)).astype(np.float32)
    wi = rng.randint(-8, 9, size=(64, 32)).astype(np.float32)
    got = ops.dense_tf32x3(torch.from_numpy(xi).to(dev), torch.from_numpy(wi).to(dev)).cpu().numpy()
    assert np.array_equal(got, xi @ wi)


def test_tf32x3_strided_views_and_untouched_neighbours(dev):
    from deep_cbrs_amar_renaissance_b200 import ops
    x, w, b, xd, wd, bd = _case(700, 128, 128, 9, dev)
    buf = torch.full((700, 512), -3.0, device=dev)
    buf[:, 128:256] = xd
    ops.dense_tf32x3(buf[:, 128:256], wd, out=buf[:, 256:384])
    want = (x.astype(np.float64) @ w.astype(np.float64)).astype(np.float32)
    assert_close(buf[:, 256:384].cpu().numpy(), want, rtol=1e-5)
    assert (buf[:, :128] == -3).all() and (buf[:, 384:] == -3).all() and torch.equal(buf[:, 128:256], xd)


def test_tf32x3_rows_do_not_depend_on_their_tile_position(dev):
    """a row partition hands the kernel slices that start anywhere: same bits as the full run"""
    from deep_cbrs_amar_renaissance_b200 import ops
    x, w, b, xd, wd, bd = _case(3000, 128, 128, 4, dev)
    full = ops.dense_tf32x3(xd, wd)
    for a, e in ((0, 3000), (1, 2999), (37, 165), (1500, 3000), (129, 130)):
        part = ops.dense_tf32x3(xd[a:e], wd)
        assert torch.equal(part, full[a:e])


def test_gcn_layer_with_the_tensor_core_transform_matches_oracle(dev, monkeypatch):
    from deep_cbrs_amar_renaissance_b200 import ops
    from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph
    from oracle import layers as ol
    adj = random_bipartite(300, 80, 6000, seed=8, n_props=30, n_links=200, dup_links=20)
    n = adj.shape[0]
    rng = np.random.RandomState(1)
    x = rng.standard_normal((n, 64)).astype(np.float32)
    w = (rng.standard_normal((64, 32)) / 8).astype(np.float32)
    b = rng.standard_normal(32).astype(np.float32)
    g = DeviceGraph.from_scipy(adj, dev)
    xd, wd, bd = (torch.from_numpy(t).to(dev) for t in (x, w, b))
    want = ol.gcn_conv(x, og.gcn_filter(adj), w, b, "relu")
    out = torch.empty(n, 32, device=dev)
    # auto: a MovieLens-sized graph stays on the FFMA kernel (round-1 bits)
    assert not ops.tf32x3_chosen(n, 64, 32)
    z_auto = ops.gcn_transform(xd, wd, n)
    assert torch.equal(z_auto, ops.dense(xd, wd))
    monkeypatch.setattr(ops, "GCN_TRANSFORM", "tf32x3")
    assert ops.tf32x3_chosen(n, 64, 32) and not ops.tf32x3_chosen(n, 40, 32) and not ops.tf32x3_chosen(n, 64, 24)
    z = ops.gcn_transform(xd, wd, n)
    ops.spmm(g.norm, z, out, bias=bd, relu=True)
    assert_close(out.cpu().numpy(), want, rtol=1e-5, what="GCN layer, tf32x3 transform")


def test_tf32x3_bf16_output_is_the_rounded_fp32_output(dev):
    """out_dtype = bf16 (the bf16-operand variant of the propagation stores Z = X W as bf16): exactly the fp32 result
    rounded to nearest even, also into a strided buffer"""
    from deep_cbrs_amar_renaissance_b200 import ops
    x, w, b, xd, wd, bd = _case(1000, 128, 128, 12, dev)
    full = ops.dense_tf32x3(xd, wd)
    assert torch.equal(ops.dense_tf32x3(xd, wd, out_dtype=torch.bfloat16), full.to(torch.bfloat16))
    buf = torch.zeros(1000, 256, device=dev, dtype=torch.bfloat16)
    ops.dense_tf32x3(xd, wd, bd, "relu", out=buf[:, 128:])
    assert torch.equal(buf[:, 128:], ops.dense_tf32x3(xd, wd, bd, "relu").to(torch.bfloat16)) and (buf[:, :128] == 0).all()
    assert ops.tf32x3_chosen(1 << 20, 128, 128, xd, buf[:, 128:]) and not ops.tf32x3_chosen(1 << 20, 128, 128, xd, buf[:, 8:136])


def test_gat_transform_on_the_tensor_cores(dev, monkeypatch):
    """z = x W with the two attention logits per row (cbrs_dense_tf32x3_attn) against float64, and a whole GAT layer
    with it against the oracle"""
    from deep_cbrs_amar_renaissance_b200 import ops
    from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph
    from oracle import layers as ol
    x, w, b, xd, wd, bd = _case(2000, 64, 32, 21, dev)
    rng = np.random.RandomState(2)
    a_s, a_n = rng.standard_normal(32).astype(np.float32), rng.standard_normal(32).astype(np.float32)
    monkeypatch.setattr(ops, "GCN_TRANSFORM", "tf32x3")
    z, p, q = ops.gat_transform(xd, wd, torch.from_numpy(a_s).to(dev), torch.from_numpy(a_n).to(dev), 2000)
    z64 = x.astype(np.float64) @ w.astype(np.float64)
    assert_close(z.cpu().numpy(), z64.astype(np.float32), rtol=1e-5, what="z")
    assert_close(p.cpu().numpy(), (z64 @ a_s).astype(np.float32), rtol=1e-5, what="p = z . a_self")
    assert_close(q.cpu().numpy(), (z64 @ a_n).astype(np.float32), rtol=1e-5, what="q = z . a_neigh")
    zf, pf, qf = ops.dense(xd, wd, rowop=2, a_self=torch.from_numpy(a_s).to(dev), a_neigh=torch.from_numpy(a_n).to(dev))
    assert float((p - pf).abs().max()) <= 5e-6 * float(pf.abs().max())
    # row slices give the same bits (row partition)
    z2, p2, q2 = ops.gat_transform(xd[100:777], wd, torch.from_numpy(a_s).to(dev), torch.from_numpy(a_n).to(dev), 2000)
    assert torch.equal(z2, z[100:777]) and torch.equal(p2, p[100:777]) and torch.equal(q2, q[100:777])
    adj = random_bipartite(300, 80, 6000, seed=8)
    n = adj.shape[0]
    xx = rng.standard_normal((n, 64)).astype(np.float32)
    g = DeviceGraph.from_scipy(adj, dev)
    zz, pp, qq = ops.gat_transform(torch.from_numpy(xx).to(dev), wd, torch.from_numpy(a_s).to(dev), torch.from_numpy(a_n).to(dev), n)
    out = torch.empty(n, 32, device=dev)
    ops.gat(g.raw, zz, pp, qq, out, bias=bd, relu=True)
    ptr, idx, _ = og.reorder_raw(adj)
    want = ol.gat_conv(xx, ptr, idx, w, a_s, a_n, b, "relu")
    assert_close(out.cpu().numpy(), want, rtol=1e-5, what="GAT layer, tensor-core transform")


def test_graphsage_dense_part_on_the_tensor_cores(dev, monkeypatch):
    """act(l2_normalize([x || agg] @ K + b)) as two 3xTF32 products with the normalisation in the second one's epilogue
    (cbrs_dense_tf32x3_ex) against float64, against the FFMA kernel, under row slices, and a whole GraphSage layer against
    the oracle"""
    from deep_cbrs_amar_renaissance_b200 import ops
    from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph
    from deep_cbrs_amar_renaissance_b200.layers import GraphSageConv
    from oracle import layers as ol
    rng = np.random.RandomState(5)
    m, f, n = 3001, 64, 48
    x = rng.standard_normal((m, f)).astype(np.float32)
    agg = rng.standard_normal((m, f)).astype(np.float32)
    agg[7] = 0.0
    x[7] = 0.0                                        # an all-zero pre-activation row when the bias is zero
    k = (rng.standard_normal((2 * f, n)) / 11).astype(np.float32)
    b = (rng.standard_normal(n) * 0.1).astype(np.float32)
    xd, ad, kd, bd = (torch.from_numpy(t).to(dev) for t in (x, agg, k, b))
    assert torch.equal(ops.sage_dense(xd, ad, kd, bd, "relu", m), ops.dense(xd, kd, bd, "relu", x2=ad, rowop=1))   # auto: FFMA
    monkeypatch.setattr(ops, "GCN_TRANSFORM", "tf32x3")
    for bias_np, bias_d in ((b, bd), (np.zeros(n, np.float32), torch.zeros(n, device=dev))):
        v = np.concatenate([x, agg], 1).astype(np.float64) @ k.astype(np.float64) + bias_np
        want = np.maximum(v / np.sqrt(np.maximum((v * v).sum(1, keepdims=True), 1e-12)), 0).astype(np.float32)
        launches = ops.LAUNCHES
        got = ops.sage_dense(xd, ad, kd, bias_d, "relu", m)
        assert ops.LAUNCHES - launches == 4               # two operand images, two tensor-core products
        assert_close(got.cpu().numpy(), want, rtol=1e-5, what="GraphSage dense part (3xTF32)")
        ffma = ops.dense(xd, kd, bias_d, "relu", x2=ad, rowop=1)
        assert float((got - ffma).abs().max()) <= 3e-6
    full = ops.sage_dense(xd, ad, kd, bd, "relu", m)
    for a, e in ((0, m), (1, 2999), (129, 300)):
        buf = torch.full((e - a, n + 16), -2.0, device=dev)
        part = ops.sage_dense(xd[a:e], ad[a:e], kd, bd, "relu", m, out=buf[:, 8:8 + n])
        assert torch.equal(part, full[a:e]) and (buf[:, :8] == -2).all() and (buf[:, 8 + n:] == -2).all()
    # a whole layer
    adj = random_bipartite(300, 80, 6000, seed=8)
    nn = adj.shape[0]
    xs = rng.standard_normal((nn, 64)).astype(np.float32)
    g = DeviceGraph.from_scipy(adj, dev)
    layer = GraphSageConv(32, activation="relu")
    layer.build([(nn, 64)])
    out = layer([torch.from_numpy(xs).to(dev), g]).cpu().numpy()
    ptr, idx, _ = og.reorder_raw(adj)
    want = ol.sage_conv(xs, ptr, idx, layer.kernel.cpu().numpy(), layer.bias.cpu().numpy(), "mean", "relu")
    assert_close(out, want, rtol=1e-5, what="GraphSage layer, tensor-core dense part")
