"""The CUDA path against golden vectors produced by the REFERENCE's own code (tests/golden/make_golden_layers.py):
reductions, attention fusion, the DGCF operator and layer, pair-list top-k.  No oracle in between."""
import os

import numpy as np
import pytest
import torch
from scipy import sparse

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "layers", "golden_layers.npz"))


@pytest.fixture(scope="module", autouse=True)
def _device():
    from deep_cbrs_amar_renaissance_b200 import ops
    ops.check_device()
    torch.cuda.set_device(0)


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("method", ["concatenation", "sum", "mean", "last", "w-sum"])
def test_reductions_vs_reference_golden(method):
    from deep_cbrs_amar_renaissance_b200.layers import ReductionLayer
    hs = [_cu(G["red_h%d" % l]) for l in range(3)]
    layer = ReductionLayer(method)
    out = layer(hs)
    if method == "w-sum":
        layer.layer.w.copy_(_cu(G["red_wsum_w"]))
        out = layer(hs)
    want = G["red_" + method]
    if method in ("concatenation", "last", "sum", "mean"):
        assert np.array_equal(out.cpu().numpy(), want), method   # same float32 additions in the same order
    else:
        assert_close(out.cpu().numpy(), want, rtol=1e-6, what=method)


@pytest.mark.parametrize("tag", ["same", "projA", "projB"])
def test_attention_fusion_vs_reference_golden(tag):
    from deep_cbrs_amar_renaissance_b200.layers import FusionLayer
    a, b = _cu(G["fus_%s_a" % tag]), _cu(G["fus_%s_b" % tag])
    layer = FusionLayer("attention")
    layer.build_for(a.shape[1], b.shape[1])
    layer.att_weight.copy_(_cu(G["fus_%s_att" % tag]))
    if "fus_%s_proj" % tag in G:
        assert layer.proj_first == bool(G["fus_%s_proj_first" % tag])
        layer.proj_weight.copy_(_cu(G["fus_%s_proj" % tag]))
    assert_close(layer([a, b]).cpu().numpy(), G["fus_%s_out" % tag], rtol=1e-5, what="attention fusion " + tag)


@pytest.mark.parametrize("tag", ["ui", "uip"])
def test_dgcf_operator_and_layer_vs_reference_golden(tag):
    from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph
    from deep_cbrs_amar_renaissance_b200.layers import DGCFConv
    n = int(G["dgcf_%s_n" % tag])
    adj = sparse.coo_matrix((G["dgcf_%s_val" % tag], (G["dgcf_%s_row" % tag], G["dgcf_%s_col" % tag])), shape=(n, n))
    g = DeviceGraph.from_scipy(adj)
    got = g.dgcf.to_scipy()
    got.sort_indices()
    assert np.array_equal(got.indptr, G["dgcf_%s_indptr" % tag]) and np.array_equal(got.indices, G["dgcf_%s_indices" % tag])
    assert_close(got.data, G["dgcf_%s_data" % tag], what="dgcf operator values")
    layer = DGCFConv(None)
    x = _cu(G["dgcf_%s_x" % tag])
    layer([x, g])
    layer.locality_adaptive.w.copy_(_cu(G["dgcf_%s_w" % tag]))
    assert_close(layer([x, g]).cpu().numpy(), G["dgcf_%s_out" % tag], what="dgcf layer")


@pytest.mark.parametrize("k", [1, 3, 5])
def test_pair_top_k_vs_reference_golden(k):
    from deep_cbrs_amar_renaissance_b200.utilities.metrics import top_k_predictions
    df = top_k_predictions(G["topk_preds"], G["topk_users"], G["topk_items"], k=k)
    assert np.array_equal(df["users"].to_numpy(), G["topk_k%d_users" % k])
    assert np.array_equal(df["items"].to_numpy(), G["topk_k%d_items" % k])   # exact ties: earlier input row first
    assert np.array_equal(df["scores"].to_numpy(), G["topk_k%d_scores" % k])
