"""Oracle vs golden vectors produced by the REFERENCE's own Two-Step / Two-Way model code run end to end
(tests/golden/make_golden_models_kg.py): step/way wiring, item and user row slices, width bookkeeping.  Host only."""
import os

import numpy as np
import pytest

from deep_cbrs_amar_renaissance_b200.data.preprocess import get_user_properties
from oracle import graph as og
from oracle import layers as ol
from tests.helpers import KG_GRAPHS, assert_close, kg_graphs, weights_struct

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "models", "golden_models_kg.npz"))
CASES = sorted({k.split("/")[0] for k in G.files if "/out/" in k})
KIND = {"GCN": "gcn", "GAT": "gat", "GraphSage": "sage", "LightGCN": "lightgcn", "DGCF": "dgcf"}
N_USERS, N_ITEMS, N_PROPS = int(G["n_users"]), int(G["n_items"]), int(G["n_props"])


def case_setup(case):
    """(kind, is two-way, constructor extras, graphs (ui, ip, up))"""
    name = case.split("-")[0]
    two_way = "TW" in name
    kind = KIND[name.replace("HybridBert", "").replace("Basic", "")[2:]]
    extra = {}
    if case.endswith("-itemconcat"):
        extra["item_node"] = "concatenation"
    if case.endswith("-uiconcat"):
        extra["user_item_node"] = "concatenation"
    if case.endswith("-mean"):
        extra.update(final_node="mean", aggregate="sum")
    tag = "sparse" if case.endswith("-sparse") else "default"
    ui, ip = kg_graphs(N_USERS, N_ITEMS, N_PROPS, **KG_GRAPHS[tag])
    up = get_user_properties(ui, ip, N_USERS, N_ITEMS)
    return kind, two_way, extra, (ui, ip, up), tag


def oracle_graph(kind, adj):
    if kind in ("gcn", "lightgcn"):
        return og.gcn_filter(adj)
    if kind == "dgcf":
        return ol.dgcf_preprocess(adj)[0]
    ptr, idx, _ = og.reorder_raw(adj)
    return (ptr, idx)


def part_weights(case, part, n_layers=2):
    """dict(embeddings=, layers=[...]) of one SequentialGNN of the model, from the golden's weight paths"""
    pre = "%s/gnn/%s/" % (case, part)
    named = {"gnn/gnn_layers/" + k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
    emb = named.pop("gnn/gnn_layers/embeddings", None)
    named["gnn/gnn_layers/embeddings"] = emb if emb is not None else np.zeros((1, 1), np.float32)
    w = weights_struct(named)
    w["layers"] += [{} for _ in range(n_layers - len(w["layers"]))]
    return dict(embeddings=emb, layers=w["layers"])


def oracle_embeddings(case):
    kind, two_way, extra, (ui, ip, up), _ = case_setup(case)
    final_node = extra.get("final_node", "concatenation")
    aggregate = extra.get("aggregate", "mean")
    step_two = dict(part_weights(case, "step_two_gnn_layers"), graph=oracle_graph(kind, ui))
    if two_way:
        one = dict(part_weights(case, "way_one_gnn_layers"), graph=oracle_graph(kind, up))
        two = dict(part_weights(case, "way_two_gnn_layers"), graph=oracle_graph(kind, ip))
        return ol.two_way(kind, one, two, step_two, N_USERS, N_ITEMS, extra.get("user_item_node", "mean"), final_node, aggregate)
    one = dict(part_weights(case, "step_one_gnn_layers"), graph=oracle_graph(kind, ip))
    return ol.two_step(kind, one, step_two, N_ITEMS, extra.get("item_node", "mean"), final_node, aggregate)


def test_all_cases_present():
    assert len(CASES) == 15 and "BasicTWDGCF-sparse" in CASES and "HybridBertTSGCN" in CASES


@pytest.mark.parametrize("tag", sorted(KG_GRAPHS))
def test_user_property_graph_is_the_reference_one(tag):
    ui, ip = kg_graphs(N_USERS, N_ITEMS, N_PROPS, **KG_GRAPHS[tag])
    up = get_user_properties(ui, ip, N_USERS, N_ITEMS)
    assert np.array_equal(up.row, G[tag + "/up_row"]) and np.array_equal(up.col, G[tag + "/up_col"])
    assert np.array_equal(up.data, G[tag + "/up_data"])


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_the_reference_models(case):
    emb = oracle_embeddings(case)
    assert_close(emb, G[case + "/out/embeddings"], rtol=2e-6, what=case + " embeddings")
    named = {k[len(case) + 1:]: G[k] for k in G.files if k.startswith(case + "/rs/")}
    named["gnn/gnn_layers/embeddings"] = np.zeros((1, 1), np.float32)
    w = weights_struct(named)
    u, i = G["u"], G["i"]
    if case.startswith("Basic"):
        scores = ol.basic_rs(emb, u, i, w["unet"], w["inet"], w["clf"])
    else:
        scores = ol.hybrid_cbrs(emb, u, i, G["ub"], G["ib"], w, feature_based=True)
    assert_close(scores, G[case + "/out/scores"], rtol=2e-6, what=case + " scores")
