"""Oracle vs golden vectors produced by the REFERENCE's own layer / metric code (tests/golden/make_golden_layers.py:
src/layers/reduction.py, fusion.py, dgcf_conv.py and src/utilities/metrics.py run unmodified over numpy-backed
tf/keras/spektral stand-ins).  Host only."""
import os

import numpy as np
import pytest
from scipy import sparse

from oracle import layers as ol
from tests.helpers import assert_close

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "layers", "golden_layers.npz"))


@pytest.mark.parametrize("method", ["concatenation", "sum", "mean", "last", "w-sum"])
def test_reductions_match_the_reference_code(method):
    hs = [G["red_h%d" % l] for l in range(3)]
    got = ol.reduce_layers(hs, method, w=G["red_wsum_w"] if method == "w-sum" else None)
    want = G["red_" + method]
    assert got.shape == want.shape
    if method == "w-sum":
        assert_close(got, want, rtol=1e-6, what=method)   # the reference sums with one reduce_sum, the oracle left to right
    else:
        assert np.array_equal(got, want), method


@pytest.mark.parametrize("tag", ["same", "projA", "projB"])
def test_attention_fusion_matches_the_reference_code(tag):
    fw = dict(att_weight=G["fus_%s_att" % tag], proj_weight=None, proj_first=None)
    if "fus_%s_proj" % tag in G:
        fw.update(proj_weight=G["fus_%s_proj" % tag], proj_first=bool(G["fus_%s_proj_first" % tag]))
    got = ol.fusion(G["fus_%s_a" % tag], G["fus_%s_b" % tag], fw)
    assert_close(got, G["fus_%s_out" % tag], rtol=1e-6, what="attention fusion " + tag)
    assert np.array_equal(ol.fusion(G["fus_cat_a"], G["fus_cat_b"]), G["fus_cat_out"])


@pytest.mark.parametrize("tag", ["ui", "uip"])
def test_dgcf_operator_and_layer_match_the_reference_code(tag):
    n = int(G["dgcf_%s_n" % tag])
    adj = sparse.coo_matrix((G["dgcf_%s_val" % tag], (G["dgcf_%s_row" % tag], G["dgcf_%s_col" % tag])), shape=(n, n))
    m, info = ol.dgcf_preprocess(adj)
    want = sparse.csr_matrix((G["dgcf_%s_data" % tag], G["dgcf_%s_indices" % tag], G["dgcf_%s_indptr" % tag]), shape=(n, n))
    assert np.array_equal(m.indptr, want.indptr) and np.array_equal(m.indices, want.indices)
    assert_close(m.data, want.data, rtol=1e-6, what="dgcf operator")   # 1-ulp class: numpy power vs 1/sqrt in gcn_filter
    out = ol.dgcf_conv(G["dgcf_%s_x" % tag], want, G["dgcf_%s_w" % tag])
    assert_close(out, G["dgcf_%s_out" % tag], rtol=1e-6, what="dgcf layer")


@pytest.mark.parametrize("k", [1, 3, 5])
def test_pair_top_k_matches_the_reference_code(k):
    preds, users, items = G["topk_preds"], G["topk_users"], G["topk_items"]
    uu, ii, ss, rows = ol.top_k_pairs(preds[:, 0].astype(np.int64), preds[:, 1].astype(np.int64), preds[:, 2], k)
    assert np.array_equal(users[uu], G["topk_k%d_users" % k])
    assert np.array_equal(items[ii - len(users)], G["topk_k%d_items" % k])   # ties: earlier input row first
    assert np.array_equal(ss, G["topk_k%d_scores" % k])
