"""Oracle vs golden vectors produced by the REFERENCE's own model code run end to end (tests/golden/make_golden_models.py):
SequentialGNN loop + reduction, family builders, embedding lookups, BasicRS / HybridCBRS wiring in every mode.  Host only."""
import os

import numpy as np
import pytest
from scipy import sparse

from oracle import graph as og
from oracle import layers as ol
from tests.helpers import assert_close, weights_struct

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "models", "golden_models.npz"))
CASES = sorted({k.split("/")[0] for k in G.files if "/" in k})
KIND = {"GCN": "gcn", "GAT": "gat", "GraphSage": "sage", "LightGCN": "lightgcn", "DGCF": "dgcf"}


def case_weights(case):
    named = {k[len(case) + 1:]: G[k] for k in G.files if k.startswith(case + "/") and "/out/" not in k and not k.endswith("proj_first")}
    pf = {k.split("/")[-2]: bool(G[k]) for k in G.files if k.startswith(case + "/") and k.endswith("proj_first")}
    return named, pf


def family(case):
    name = case.split("-")[0].replace("HybridBert", "").replace("Basic", "")
    return KIND[name]


def adjacency():
    n = int(G["n_nodes"])
    return sparse.coo_matrix((G["adj_data"], (G["adj_row"], G["adj_col"])), shape=(n, n))


def oracle_graph(kind, adj):
    if kind in ("gcn", "lightgcn"):
        return og.gcn_filter(adj)
    if kind == "dgcf":
        return ol.dgcf_preprocess(adj)[0]
    ptr, idx, _ = og.reorder_raw(adj)
    return (ptr, idx)


def test_all_cases_present():
    assert len(CASES) == 11 and "BasicDGCF" in CASES and "HybridBertGAT-attention" in CASES


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_the_reference_models(case):
    named, pf = case_weights(case)
    w = weights_struct(named, pf)
    kind = family(case)
    n_layers = 2
    w["layers"] += [{} for _ in range(n_layers - len(w["layers"]))]
    adj = adjacency()
    emb = ol.propagate(kind, w["embeddings"], oracle_graph(kind, adj), w["layers"])
    assert_close(emb, G[case + "/out/embeddings"], rtol=2e-6, what=case + " embeddings")
    u, i = G["u"], G["i"]
    if case.startswith("Basic"):
        scores = ol.basic_rs(emb, u, i, w["unet"], w["inet"], w["clf"])
    else:
        scores = ol.hybrid_cbrs(emb, u, i, G["ub"], G["ib"], w, feature_based="entity" not in case)
    assert_close(scores, G[case + "/out/scores"], rtol=2e-6, what=case + " scores")
