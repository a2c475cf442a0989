"""bf16 operand storage for the sparse kernel (BASELINE config 5's "bf16 feature variant"): the stored transform
Z = X W is rounded to bf16, products and sums stay fp32.

Exactness claims tested here: (1) the bf16 sparse kernel equals the fp32 kernel run on the same values widened to
float32, bit for bit (bf16 -> fp32 is exact, the accumulation order is the same); (2) the dense kernel's bf16 output
equals its fp32 output rounded to nearest even.  Tolerances: a model with bf16 operands vs the oracle that rounds Z
the same way 1e-4 (an element whose fp32 value sits within 1e-7 of a bf16 rounding boundary may round the other way),
vs the fp32 oracle 1e-2 (bf16 has 8 significant bits: 2^-9 = 2e-3 per stored element)."""
import numpy as np
import pytest
import torch

from oracle import graph as og
from oracle import layers as ol
from tests.helpers import assert_close, export_weights, random_bipartite

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device():
    from deep_cbrs_amar_renaissance_b200 import ops
    ops.check_device()
    torch.cuda.set_device(0)


@pytest.mark.parametrize("d", [128, 64, 16, 12, 7, 1])
def test_bf16_spmm_equals_fp32_spmm_on_widened_values(d):
    from deep_cbrs_amar_renaissance_b200 import ops
    from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph
    adj = random_bipartite(400, 30, 9000, seed=d)  # 30 hot items: rows longer than a 64-edge chunk -> heavy-row path
    g = DeviceGraph.from_scipy(adj, chunk_edges=64)
    csr = g.norm
    assert csr.chunks["n_heavy"] > 0
    rng = np.random.RandomState(d)
    x16 = torch.from_numpy(rng.standard_normal((430, d)).astype(np.float32)).cuda().to(torch.bfloat16)
    bias = torch.from_numpy(rng.standard_normal(d).astype(np.float32)).cuda()
    a = torch.empty(430, d, device="cuda")
    b = torch.empty(430, d, device="cuda")
    ops.spmm(csr, x16, a, bias=bias, relu=True)
    ops.spmm(csr, x16.float(), b, bias=bias, relu=True)
    assert torch.equal(a, b)
    for agg in (1, 2):  # sum / mean over the raw edge list
        ops.spmm(g.raw, x16, a, agg=agg)
        ops.spmm(g.raw, x16.float(), b, agg=agg)
        assert torch.equal(a, b)


@pytest.mark.parametrize("m,k,n", [(1000, 128, 128), (333, 16, 16), (257, 48, 20), (64, 32, 64)])
def test_dense_bf16_output_is_the_rounded_fp32_output(m, k, n):
    from deep_cbrs_amar_renaissance_b200 import ops
    rng = np.random.RandomState(m)
    x = torch.from_numpy(rng.standard_normal((m, k)).astype(np.float32)).cuda()
    w = torch.from_numpy(rng.standard_normal((k, n)).astype(np.float32)).cuda()
    f = ops.dense(x, w)
    h = ops.dense(x, w, out_dtype=torch.bfloat16)
    assert h.dtype == torch.bfloat16 and torch.equal(h, f.to(torch.bfloat16))
    assert np.array_equal(ol.bf16_round(f.cpu().numpy()), h.float().cpu().numpy())  # the oracle's rounding helper agrees


def test_gcn_with_bf16_operands_matches_the_rounding_oracle():
    from deep_cbrs_amar_renaissance_b200.keras_like import set_seed
    from deep_cbrs_amar_renaissance_b200.models import basic
    set_seed(3)
    n_users, n_items = 500, 300
    adj = random_bipartite(n_users, n_items, 12000, seed=8)
    model = basic.BasicGCN(adj, n_hiddens=[32, 32, 32], embedding_dim=32, dense_units=[48, 48], clf_units=[64, 64])
    rng = np.random.RandomState(0)
    u = rng.randint(0, n_users, 512)
    i = rng.randint(0, n_items, 512) + n_users
    model((u, i))
    w = export_weights(model)
    a_hat = og.gcn_filter(adj)
    fp32 = model.gnn(None).cpu().numpy().copy()
    model.gnn.gnn_layers.set_feature_dtype("bf16")
    got = model.gnn(None).cpu().numpy()
    want16 = ol.propagate("gcn", w["embeddings"], a_hat, w["layers"], operand_dtype="bf16")
    want32 = ol.propagate("gcn", w["embeddings"], a_hat, w["layers"])
    assert_close(got, want16, rtol=1e-4, what="bf16 operands vs rounding oracle")
    assert_close(got, want32, rtol=1e-2, what="bf16 operands vs fp32 oracle")
    assert_close(fp32, want32, what="fp32 path untouched")
    assert not np.array_equal(got, fp32)
    scores = model((u, i)).cpu().numpy()
    ref = ol.basic_rs(want32, u, i, w["unet"], w["inet"], w["clf"])
    assert_close(scores, ref, rtol=1e-2, what="scores with bf16 operands")
    with pytest.raises(NotImplementedError):
        basic.BasicGAT(adj, n_hiddens=[8], embedding_dim=8).gnn.gnn_layers.set_feature_dtype("bf16")
