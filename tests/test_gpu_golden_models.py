"""The CUDA path against golden vectors produced by the REFERENCE's own model code run end to end
(tests/golden/make_golden_models.py): same constructor keywords, same weights (loaded by path), same batch ->
same propagated embeddings and scores within 1e-5 (hybrid 2e-5).  No oracle in between."""
import os

import numpy as np
import pytest
import torch
from scipy import sparse

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "models", "golden_models.npz"))
CASES = sorted({k.split("/")[0] for k in G.files if "/" in k})
COMMON = dict(n_hiddens=[8, 8], n_layers=2, embedding_dim=8, l2_regularizer=1e-4, aggregate="mean", dropout_rate=0.0,
              final_node="concatenation", activation="relu")
EXTRA = {
    "BasicGCN-notowers": dict(dense_units=[], clf_units=[16]),
    "HybridBertGCN-feature": dict(feature_based=True),
    "HybridBertGCN-entity": dict(feature_based=False),
    "HybridBertGAT-attention": dict(feature_based=True, fusion_method="attention"),
    "HybridBertLightGCN-residual": dict(feature_based=True, residual=True),
    "HybridBertGraphSage-entity-attention": dict(feature_based=False, fusion_method="attention"),
}


@pytest.mark.parametrize("case", CASES)
def test_models_reproduce_the_reference_run(case):
    from deep_cbrs_amar_renaissance_b200 import ops
    from deep_cbrs_amar_renaissance_b200.models import basic, hybrid
    ops.check_device()
    torch.cuda.set_device(0)
    n = int(G["n_nodes"])
    adj = sparse.coo_matrix((G["adj_data"], (G["adj_row"], G["adj_col"])), shape=(n, n))
    is_hybrid = case.startswith("Hybrid")
    kw = dict(COMMON)
    kw.update(dict(dense_units=[[12, 12], [14, 6], [16, 16]], clf_units=[16, 16]) if is_hybrid else dict(dense_units=[12, 12], clf_units=[16, 16]))
    kw.update(EXTRA.get(case, {}))
    model = getattr(hybrid if is_hybrid else basic, case.split("-")[0])(adj, **kw)
    inputs = (G["u"], G["i"], G["ub"], G["ib"]) if is_hybrid else (G["u"], G["i"])
    model(inputs)   # creates the weights (experiment.py:166)
    names = [nm for nm, _ in model.named_weights()]
    golden_names = sorted(k[len(case) + 1:] for k in G.files if k.startswith(case + "/") and "/out/" not in k and not k.endswith("proj_first"))
    assert sorted(names) == golden_names, (sorted(set(names) ^ set(golden_names)))
    for nm, w in model.named_weights():
        w.copy_(torch.from_numpy(np.ascontiguousarray(G[case + "/" + nm], dtype=np.float32)).reshape(w.shape).cuda())
    for k in G.files:   # attention fusions that project: same side projected as in the reference run
        if k.startswith(case + "/") and k.endswith("proj_first"):
            assert getattr(model.rs, k.split("/")[-2]).proj_first == bool(G[k])
    emb = model.gnn(None).cpu().numpy()
    assert_close(emb, G[case + "/out/embeddings"], rtol=1e-5, what=case + " embeddings")
    scores = model(inputs).cpu().numpy()
    assert_close(scores, G[case + "/out/scores"], rtol=2e-5 if is_hybrid else 1e-5, what=case + " scores")
