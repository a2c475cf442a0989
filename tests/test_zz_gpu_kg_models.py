"""Two-Step / Two-Way variants (scope row (f)-4) on the CUDA path against golden vectors produced by the REFERENCE's
own tsgnn.py / twgnn.py run end to end (tests/golden/make_golden_models_kg.py): same constructor call, same weights
(loaded by path), same batch -> same propagated embeddings and scores within 1e-5 (hybrid 2e-5).
(File name sorts last on purpose: these cases were added after the round's last GPU session.)"""
import os

import numpy as np
import pytest
import torch

from tests.helpers import KG_GRAPHS, assert_close, kg_graphs, kg_model

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "models", "golden_models_kg.npz"))
CASES = sorted({k.split("/")[0] for k in G.files if "/out/" in k})
N_USERS, N_ITEMS, N_PROPS = int(G["n_users"]), int(G["n_items"]), int(G["n_props"])


def build(case):
    from deep_cbrs_amar_renaissance_b200 import ops
    from deep_cbrs_amar_renaissance_b200.data.preprocess import get_user_properties
    ops.check_device()
    torch.cuda.set_device(0)
    ui, ip = kg_graphs(N_USERS, N_ITEMS, N_PROPS, **KG_GRAPHS["sparse" if case.endswith("-sparse") else "default"])
    graphs = (ui, ip, get_user_properties(ui, ip, N_USERS, N_ITEMS))
    model, _ = kg_model(case, graphs, N_USERS, N_ITEMS)
    is_hybrid = case.startswith("Hybrid")
    inputs = (G["u"], G["i"], G["ub"], G["ib"]) if is_hybrid else (G["u"], G["i"])
    model(inputs)   # creates the weights (experiment.py:166)
    names = sorted(nm for nm, _ in model.named_weights())
    golden_names = sorted(k[len(case) + 1:] for k in G.files if k.startswith(case + "/") and "/out/" not in k)
    assert names == golden_names, sorted(set(names) ^ set(golden_names))
    for nm, w in model.named_weights():
        w.copy_(torch.from_numpy(np.ascontiguousarray(G[case + "/" + nm], dtype=np.float32)).reshape(w.shape).cuda())
    return model, inputs, is_hybrid


@pytest.mark.parametrize("case", CASES)
def test_kg_models_reproduce_the_reference_run(case):
    model, inputs, is_hybrid = build(case)
    emb = model.gnn(None).cpu().numpy()
    assert_close(emb, G[case + "/out/embeddings"], rtol=1e-5, what=case + " embeddings")
    scores = model(inputs).cpu().numpy()
    assert_close(scores, G[case + "/out/scores"], rtol=2e-5 if is_hybrid else 1e-5, what=case + " scores")
    # a second call reuses the concatenation buffers of every step: same result
    assert np.array_equal(model(inputs).cpu().numpy(), scores)


def test_user_property_graph_normalisation_on_device():
    """the user-property adjacency arrives as float64 ones (coo_matrix of a dense float64 array); the device build
    takes it as float32 and its A_hat equals the oracle's gcn_filter bit for bit in structure, 1e-6 in value"""
    from deep_cbrs_amar_renaissance_b200.data.preprocess import get_user_properties
    from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph
    from oracle import graph as og
    ui, ip = kg_graphs(N_USERS, N_ITEMS, N_PROPS)
    up = get_user_properties(ui, ip, N_USERS, N_ITEMS)
    got = DeviceGraph.from_scipy(up).norm.to_scipy()
    want = og.gcn_filter(up)
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert_close(got.data, want.data, rtol=1e-6, what="A_hat of the user-property graph")
