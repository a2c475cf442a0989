"""cbrs_dense_tc_bf16: the tensor-core Dense layer over bf16-STORED sources, fed by TMA (tiled boxes for consecutive rows,
tile::gather4 for indexed rows), against (a) the oracle that rounds the operands to bf16 and sums in float64 (1e-4 of the
output scale, as for cbrs_dense_tc) and (b) cbrs_dense_tc itself on the fp32 table: both kernels multiply the same bf16
operands in the same K order, so they must agree to fp32 summation noise (2e-6 of the scale stated; bit-equal in practice).
cbrs_convert_f32_bf16 is checked bit for bit against the oracle's round-to-nearest-even."""
import numpy as np
import pytest
import torch

from oracle import layers as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device():
    from deep_cbrs_amar_renaissance_b200 import ops
    ops.check_device()
    torch.cuda.set_device(0)


def oracle_dense_bf16(a, w, b, act):
    y = ol.bf16_round(a).astype(np.float64) @ ol.bf16_round(w).astype(np.float64)
    if b is not None:
        y = y + b.astype(np.float64)
    return ol.activation(act)(y.astype(np.float32))


def cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def close(got, want, what, rtol=1e-4):
    scale = max(np.abs(want).max(), 1e-30)
    err = np.abs(got.astype(np.float64) - want.astype(np.float64)).max()
    assert err <= rtol * scale, "{}: max abs err {:.3e} (scale {:.3e})".format(what, err, scale)


def test_convert_is_round_to_nearest_even():
    from deep_cbrs_amar_renaissance_b200 import ops
    rng = np.random.RandomState(0)
    x = rng.standard_normal((333, 72)).astype(np.float32)
    x[0, :8] = [1.0, 1.00390625, 1.01171875, -1.00390625, 3.0e38, 1e-40, 0.0, -0.0]   # ties, large, subnormal, zeros
    buf = cuda(np.zeros((333, 80), np.float32))
    buf[:, 4:76] = cuda(x)
    got = ops.to_bf16(buf[:, 4:76])      # strided source view
    assert got.dtype == torch.bfloat16 and tuple(got.shape) == (333, 72)
    want = ol.bf16_round(x)
    assert np.array_equal(got.float().cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("m,k,n,act,bias", [(300, 768, 256, "relu", True), (128, 64, 64, None, True), (77, 128, 16, "tanh", False),
                                            (1000, 192, 40, None, True), (5, 256, 8, "sigmoid", True), (4097, 128, 128, "relu", True),
                                            (148 * 256 * 3 + 17, 768, 256, "relu", True),     # 3-4 tiles per CTA, one accumulator set
                                            (148 * 256 * 4 + 300, 256, 64, "relu", True)])    # 4-5 tiles per CTA, two sets
def test_consecutive_rows(m, k, n, act, bias):
    from deep_cbrs_amar_renaissance_b200 import ops
    rng = np.random.RandomState((m + k + n) % 65536)
    a = rng.standard_normal((m, k)).astype(np.float32)
    w = (rng.standard_normal((k, n)) / np.sqrt(k)).astype(np.float32)
    b = (rng.standard_normal(n) * 0.1).astype(np.float32) if bias else None
    a_d, w_d, b_d = cuda(a), cuda(w), (cuda(b) if bias else None)
    got = ops.dense_tc_bf16(ops.to_bf16(a_d), w_d, b_d, act)
    assert got.shape == (m, n) and got.dtype == torch.float32
    ref_tc = ops.dense_tc(a_d, w_d, b_d, act)
    close(got.cpu().numpy(), ref_tc.cpu().numpy(), "vs cbrs_dense_tc %dx%dx%d" % (m, k, n), rtol=2e-6)
    if m <= 5000:
        close(got.cpu().numpy(), oracle_dense_bf16(a, w, b, act), "vs oracle %dx%dx%d %s" % (m, k, n, act))
    else:   # the float64 oracle on a row sample of every tile position
        rows = np.unique(np.concatenate([np.arange(0, m, 997), np.arange(m - 300, m)]))
        close(got.cpu().numpy()[rows], oracle_dense_bf16(a[rows], w, b, act), "vs oracle (sampled rows) %dx%dx%d" % (m, k, n))


def test_gathered_two_sources_strided_views_and_bf16_output():
    """the scorer's lookup + concat: two indexed bf16 tables (one a column slice), strided output view, duplicate ids"""
    from deep_cbrs_amar_renaissance_b200 import ops
    rng = np.random.RandomState(1)
    t1 = rng.standard_normal((500, 128 + 16)).astype(np.float32)
    t2 = rng.standard_normal((400, 64)).astype(np.float32)
    m = 1000
    i1 = rng.randint(0, 500, size=m)
    i1[:64] = i1[0]
    i2 = rng.randint(0, 400, size=m)
    w = (rng.standard_normal((192, 64)) / 14).astype(np.float32)
    b = (rng.standard_normal(64) * 0.1).astype(np.float32)
    src1 = ops.to_bf16(cuda(t1))[:, 8:136]          # ld 144, offset 16 bytes
    src2 = ops.to_bf16(cuda(t2))
    buf = torch.zeros(m, 72, device="cuda")
    out = ops.dense_tc_bf16(src1, cuda(w), cuda(b), "sigmoid", x2=src2, idx1=cuda(i1), idx2=cuda(i2), out=buf[:, 4:68])
    a = np.concatenate([t1[i1][:, 8:136], t2[i2]], axis=1)
    want = oracle_dense_bf16(a, w, b, "sigmoid")
    close(out.cpu().numpy(), want, "gather + concat")
    assert float(buf[:, :4].abs().max()) == 0.0 and float(buf[:, 68:].abs().max()) == 0.0   # nothing written outside the view
    # one indexed and one consecutive source; bf16 output = the fp32 output rounded to nearest even
    f32 = ops.dense_tc_bf16(src1, cuda(w), cuda(b), "relu", x2=ops.to_bf16(cuda(t2[i2])), idx1=cuda(i1))
    b16 = ops.dense_tc_bf16(src1, cuda(w), cuda(b), "relu", x2=ops.to_bf16(cuda(t2[i2])), idx1=cuda(i1), out_dtype=torch.bfloat16)
    close(f32.cpu().numpy(), oracle_dense_bf16(a, w, b, "relu"), "gathered + consecutive")
    assert b16.dtype == torch.bfloat16
    assert np.array_equal(b16.float().cpu().numpy().view(np.uint32), ol.bf16_round(f32.cpu().numpy()).view(np.uint32))
    # a prepared image can be reused across calls
    image = ops.dense_tc_image(cuda(w))
    again = ops.dense_tc_bf16(src1, cuda(w), cuda(b), "sigmoid", x2=src2, idx1=cuda(i1), idx2=cuda(i2), image=image)
    assert torch.equal(again, out.contiguous())


def test_large_gather_matches_the_fp32_table_kernel():
    """tower shape, many tiles per CTA, random ids over a table larger than the batch"""
    from deep_cbrs_amar_renaissance_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    table = torch.randn(50000, 768, device="cuda", generator=g) * 0.5
    idx = torch.randint(0, 50000, (148 * 256 * 2 + 77,), device="cuda", generator=g)
    w = torch.randn(768, 256, device="cuda", generator=g) / 768 ** 0.5
    b = torch.randn(256, device="cuda", generator=g) * 0.1
    got = ops.dense_tc_bf16(ops.to_bf16(table), w, b, "relu", idx1=idx)
    ref = ops.dense_tc(table, w, b, "relu", idx1=idx)
    close(got.cpu().numpy(), ref.cpu().numpy(), "gathered 768 -> 256 vs cbrs_dense_tc", rtol=2e-6)


def test_refuses_what_it_cannot_take():
    from deep_cbrs_amar_renaissance_b200 import _lib, ops
    x = torch.zeros(16, 96, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.CbrsError, match="multiples of 64"):
        ops.dense_tc_bf16(x, torch.zeros(96, 16, device="cuda"))
    with pytest.raises(_lib.CbrsError, match="stored as bf16"):
        ops.dense_tc_bf16(torch.zeros(16, 64, device="cuda"), torch.zeros(64, 16, device="cuda"))
    assert ops.dense_tc_bf16_eligible(768, 0, 256) and not ops.dense_tc_bf16_eligible(96, 0, 16)


def test_hybrid_model_with_a_bf16_content_table():
    """HybridBertGCN with the content table stored as bf16 (TMA-fed towers) == the same model reading the fp32 table through
    cbrs_dense_tc (same operands after rounding), pair scores and catalog top-k; the fp32 scorer refuses a bf16 table"""
    from deep_cbrs_amar_renaissance_b200.keras_like import set_seed
    from deep_cbrs_amar_renaissance_b200.models import hybrid
    from tests.helpers import random_bipartite
    from tests.test_gpu_models import _randomise
    set_seed(42)
    n_users, n_items, b = 300, 200, 2048
    adj = random_bipartite(n_users, n_items, 6000, seed=7)
    model = hybrid.HybridBertGCN(adj, embedding_dim=16, n_hiddens=[16, 16], dense_units=[[48, 48], [256, 64], [64, 64]],
                                 clf_units=[64, 64], feature_based=True, l2_regularizer=1e-4)
    rng = np.random.RandomState(3)
    table = (rng.standard_normal((n_users + n_items, 768)) * 0.5).astype(np.float32)
    u = rng.randint(0, n_users, size=b)
    i = rng.randint(0, n_items, size=b) + n_users
    model.set_content_table(table)
    model((u, i))
    _randomise(model, seed=4)
    model.set_scorer_precision("bf16")
    want = model((u, i)).cpu().numpy()
    want_ids, want_scores = model.recommend_top_k(n_users, n_items, k=10, precision="bf16")
    model.set_content_table(table, dtype="bf16")
    assert model.content_table.dtype == torch.bfloat16
    got = model((u, i)).cpu().numpy()
    close(got, want, "pair scores, bf16 table vs fp32 table", rtol=2e-6)
    ids, scores = model.recommend_top_k(n_users, n_items, k=10, precision="bf16")
    close(scores.cpu().numpy(), want_scores.cpu().numpy(), "catalog scores", rtol=2e-6)
    assert (ids == want_ids).float().mean().item() > 0.999
    model.set_scorer_precision("fp32")
    with pytest.raises(ValueError, match="set_scorer_precision"):
        model((u, i))
