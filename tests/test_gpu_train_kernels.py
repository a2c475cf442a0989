"""Kernel-level parity of csrc/train.cu through the C ABI: each backward / optimiser kernel against numpy
(float64) on seeded inputs, including ragged shapes, gathered sources, duplicate indices and empty rows."""
import numpy as np
import pytest
import torch

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device():
    from deep_cbrs_amar_renaissance_b200 import ops
    ops.check_device()
    torch.cuda.set_device(0)


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("m,f1,f2,n", [(1024, 96, 0, 64), (1000, 48, 48, 33), (9746, 16, 0, 16), (37, 5, 3, 7),
                                       (70000, 32, 32, 16), (512, 768, 0, 256)])
def test_dense_grad_w_matches_numpy(m, f1, f2, n):
    from deep_cbrs_amar_renaissance_b200 import ops
    rng = np.random.RandomState(m + n)
    t1 = rng.standard_normal((m + 11, f1)).astype(np.float32)
    idx1 = rng.randint(0, m + 11, size=m)
    x2 = rng.standard_normal((m, f2)).astype(np.float32) if f2 else None
    dpre = rng.standard_normal((m, n)).astype(np.float32)
    dw, db = ops.dense_grad_w(_cu(t1), _cu(dpre), x2=_cu(x2) if f2 else None, idx1=_cu(idx1.astype(np.int64)))
    a = t1[idx1].astype(np.float64)
    if f2:
        a = np.concatenate([a, x2.astype(np.float64)], axis=1)
    assert_close(dw.cpu().numpy(), a.T @ dpre.astype(np.float64), rtol=2e-6 * np.sqrt(m), what="dW")
    assert_close(db.cpu().numpy(), dpre.astype(np.float64).sum(0), rtol=2e-6 * np.sqrt(m), what="db")
    # fixed reduction tree: bit-identical on a second run
    dw2, db2 = ops.dense_grad_w(_cu(t1), _cu(dpre), x2=_cu(x2) if f2 else None, idx1=_cu(idx1.astype(np.int64)))
    assert torch.equal(dw, dw2) and torch.equal(db, db2)


def test_colsum_and_strided_views():
    from deep_cbrs_amar_renaissance_b200 import ops
    rng = np.random.RandomState(0)
    big = rng.standard_normal((3000, 48)).astype(np.float32)
    t = _cu(big)
    view = t[:, 16:32]
    assert_close(ops.colsum(view).cpu().numpy(), big[:, 16:32].astype(np.float64).sum(0), rtol=1e-5, what="colsum")
    out = ops.axpby(view, 2.0, t[:, 32:48], -0.5)
    assert_close(out.cpu().numpy(), 2.0 * big[:, 16:32] - 0.5 * big[:, 32:48], rtol=1e-6, what="axpby")
    assert_close(ops.axpby(view, 0.25).cpu().numpy(), 0.25 * big[:, 16:32], rtol=1e-6, what="scale")


@pytest.mark.parametrize("act", ["relu", "sigmoid", "tanh", None])
def test_act_grad(act):
    from deep_cbrs_amar_renaissance_b200 import ops
    rng = np.random.RandomState(1)
    pre = rng.standard_normal((257, 19)).astype(np.float32)
    out = {"relu": np.maximum(pre, 0), "sigmoid": 1 / (1 + np.exp(-pre)), "tanh": np.tanh(pre), None: pre}[act].astype(np.float32)
    g = rng.standard_normal(pre.shape).astype(np.float32)
    want = {"relu": g * (out > 0), "sigmoid": g * out * (1 - out), "tanh": g * (1 - out * out), None: g}[act]
    assert_close(ops.act_grad(_cu(g), _cu(out), act).cpu().numpy(), want, rtol=1e-6, what="act_grad")


def test_transpose():
    from deep_cbrs_amar_renaissance_b200 import ops
    a = np.arange(67 * 130, dtype=np.float32).reshape(67, 130)
    assert np.array_equal(ops.transpose(_cu(a)).cpu().numpy(), a.T)


def test_scatter_add_rows_with_duplicates_is_deterministic():
    from deep_cbrs_amar_renaissance_b200 import ops
    rng = np.random.RandomState(2)
    m, d, n = 5000, 48, 700
    idx = rng.randint(0, n, size=m)
    idx[:600] = 13  # one very hot row
    src = rng.standard_normal((m, d)).astype(np.float32)
    base = rng.standard_normal((n, d)).astype(np.float32)
    dst = _cu(base.copy())
    ops.scatter_add_rows(_cu(src), _cu(idx.astype(np.int64)), dst)
    # the kernel's order: ascending m inside each run, float32 accumulation starting from the table row
    want = base.copy()
    for r in range(m):
        want[idx[r]] += src[r]
    assert np.array_equal(dst.cpu().numpy(), want)
    dst2 = _cu(base.copy())
    ops.scatter_add_rows(_cu(src), _cu(idx.astype(np.int64)), dst2)
    assert torch.equal(dst, dst2)


def test_l2norm_forward_and_backward_match_autograd():
    from deep_cbrs_amar_renaissance_b200 import ops
    rng = np.random.RandomState(3)
    v = rng.standard_normal((333, 24)).astype(np.float32)
    v[5] = 0.0  # the 1e-12 floor
    g = rng.standard_normal(v.shape).astype(np.float32)
    for relu in (True, False):
        vt = torch.tensor(v, dtype=torch.float64, requires_grad=True)
        n = vt / torch.sqrt(torch.clamp((vt * vt).sum(1, keepdim=True), min=1e-12))
        out = torch.relu(n) if relu else n
        out.backward(torch.tensor(g, dtype=torch.float64))
        got_out = ops.l2norm_act(_cu(v), relu=relu)
        assert_close(got_out.cpu().numpy(), out.detach().numpy(), rtol=1e-6, what="l2norm_act")
        got = ops.l2norm_relu_grad(_cu(v), _cu(g), relu=relu)
        rows = np.arange(len(v)) != 5
        assert_close(got.cpu().numpy()[rows], vt.grad.numpy()[rows], rtol=1e-5, what="l2norm grad")
        assert_close(got.cpu().numpy()[5], vt.grad.numpy()[5], rtol=1e-5, what="l2norm grad (floored row)")


def test_scale_rows_inv_degree():
    from deep_cbrs_amar_renaissance_b200 import ops
    deg = np.array([3, 0, 1, 7, 0, 2], dtype=np.int64)
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    x = np.random.RandomState(4).standard_normal((6, 10)).astype(np.float32)
    want = np.where(deg[:, None] > 0, x / np.maximum(deg, 1)[:, None].astype(np.float32), 0)
    assert np.array_equal(ops.scale_rows_inv_degree(_cu(x), _cu(rowptr)).cpu().numpy(), want.astype(np.float32))


def test_bce_matches_keras_formula():
    from deep_cbrs_amar_renaissance_b200 import ops
    rng = np.random.RandomState(5)
    p = rng.uniform(0, 1, size=1500).astype(np.float32)
    p[:4] = [0.0, 1.0, 1e-9, 1 - 1e-9]  # inside the clip: no gradient
    y = rng.randint(0, 2, size=1500).astype(np.float32)
    loss, dp, correct = ops.bce(_cu(p), _cu(y))
    # Keras clips in the tensor's dtype: float32(1 - 1e-7) = 1 - 1.19e-7, so that is the upper bound (and 1 - pc is
    # evaluated in float32 as well)
    lo, hi = np.float32(1e-7), np.float32(1) - np.float32(1e-7)
    pc32 = np.clip(p, lo, hi)
    pc = pc32.astype(np.float64)
    one_minus = (np.float32(1) - pc32).astype(np.float64)
    want = -(y * np.log(pc) + (1 - y) * np.log(one_minus)).mean()
    assert abs(float(loss.item()) - want) <= 1e-5 * abs(want)
    inside = (p >= lo) & (p <= hi)
    dwant = np.where(inside, (-(y / pc) + (1 - y) / one_minus) / len(p), 0.0)
    assert_close(dp.cpu().numpy(), dwant, rtol=1e-5, what="dL/dp")
    assert int(correct.item()) == int(((p > 0.5) == (y > 0.5)).sum())


def test_adam_single_and_multi_agree_with_formula():
    from deep_cbrs_amar_renaissance_b200 import ops
    from oracle import train as ot
    rng = np.random.RandomState(6)
    shapes = [(300, 16), (16,), (48, 64), (1,), (100000,)] * 11  # 55 tensors: more than one launch of the multi kernel
    ws = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    gs = [rng.standard_normal(s).astype(np.float32) * 0.1 for s in shapes]
    l2s = [1e-4 if k % 2 == 0 else 0.0 for k in range(len(shapes))]
    w1 = [_cu(w.copy()) for w in ws]
    w2 = [_cu(w.copy()) for w in ws]
    m1 = [torch.zeros_like(w) for w in w1]; v1 = [torch.zeros_like(w) for w in w1]
    m2 = [torch.zeros_like(w) for w in w2]; v2 = [torch.zeros_like(w) for w in w2]
    g = [_cu(x) for x in gs]
    lr_dev = torch.zeros(1, device="cuda")
    ref = [(w.astype(np.float64), np.zeros(w.shape), np.zeros(w.shape)) for w in ws]
    for t in (1, 2, 3):
        lr_t = 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        for k in range(len(ws)):
            ops.adam_step(w1[k], g[k], m1[k], v1[k], lr_t, 0.9, 0.999, 1e-7, l2s[k])
        lr_dev.fill_(lr_t)  # the CUDA-graph form: rate read from device memory
        ops.adam_step_multi(w2, g, m2, v2, l2s, 0.0, 0.9, 0.999, 1e-7, lr_t_dev=lr_dev)
        ref = [ot.adam_update(w, gs[k].astype(np.float64) + 2 * l2s[k] * w, m, v, t, lr=1e-3) for k, (w, m, v) in enumerate(ref)]
    for k in range(len(ws)):
        assert torch.equal(w1[k], w2[k])
        assert_close(w1[k].cpu().numpy(), ref[k][0], atol=1e-6, what="adam tensor %d" % k)


def test_sum_squares():
    from deep_cbrs_amar_renaissance_b200 import ops
    w = np.random.RandomState(7).standard_normal(70001).astype(np.float32)
    out = torch.full((1,), 2.0, device="cuda")
    ops.sum_squares(_cu(w), 1e-2, out, accumulate=True)
    want = 2.0 + 1e-2 * (w.astype(np.float64) ** 2).sum()
    assert abs(float(out.item()) - want) <= 1e-5 * want
    out2 = torch.full((1,), 2.0, device="cuda")
    ops.sum_squares(_cu(w), 1e-2, out2, accumulate=True)
    assert torch.equal(out, out2)  # fixed two-level order: reproducible
    big = np.random.RandomState(8).standard_normal(3_000_001).astype(np.float32)  # several CTAs
    o3 = torch.zeros(1, device="cuda")
    ops.sum_squares(_cu(big), 1.0, o3, accumulate=False)
    assert abs(float(o3.item()) - (big.astype(np.float64) ** 2).sum()) <= 1e-5 * 3e6
