"""Size-independent properties of the host-side pieces added for the Two-Step / Two-Way variants and the experiment
driver, over hypothesis-generated inputs (host only, a few seconds)."""
import itertools

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st
from scipy import sparse

from deep_cbrs_amar_renaissance_b200.data.preprocess import build_adjacency_matrix, get_user_properties
from deep_cbrs_amar_renaissance_b200.utilities.utils import make_grid, nested_dict_update
from oracle import graph as og


@st.composite
def kg_inputs(draw):
    n_users, n_items, n_props = draw(st.integers(1, 9)), draw(st.integers(1, 8)), draw(st.integers(1, 7))
    seed = draw(st.integers(0, 10 ** 6))
    rng = np.random.RandomState(seed)
    n_r, n_t = draw(st.integers(0, 40)), draw(st.integers(0, 30))
    ratings = np.stack([rng.randint(0, n_users, n_r), rng.randint(0, n_items, n_r) + n_users, rng.randint(0, 2, n_r)], axis=1)
    triples = np.stack([rng.randint(0, n_items, n_t), rng.randint(0, n_props, n_t) + n_items, np.ones(n_t, np.int64)], axis=1)
    return n_users, n_items, n_props, ratings.astype(np.int64), triples.astype(np.int64), draw(st.booleans())


@settings(max_examples=40, deadline=None)
@given(kg_inputs())
def test_user_property_graph_equals_the_dense_recipe(inp):
    """all-sparse get_user_properties == the reference's dense recipe (restated in oracle.graph.user_properties) for
    empty, ragged, duplicated and directed inputs; u-p are linked iff some item is linked to both"""
    n_users, n_items, n_props, ratings, triples, sym = inp
    ui, ip = build_adjacency_matrix(ratings, np.arange(n_users), np.arange(n_items), props_triples=triples,
                                    props=np.arange(n_props), type_adjacency='unary-kg', symmetric_adjacency=sym)
    got, want = get_user_properties(ui, ip, n_users, n_items), og.user_properties(ui, ip, n_users, n_items)
    assert got.shape == want.shape == (n_users + n_props, n_users + n_props) and got.dtype == want.dtype
    assert np.array_equal(got.row, want.row) and np.array_equal(got.col, want.col) and np.array_equal(got.data, want.data)
    if sym:
        likes = {(u, i - n_users) for u, i, y in ratings if y == 1}
        has = {(i, p - n_items) for i, p, _ in triples}
        expect = {(u, p) for (u, i) in likes for (j, p) in has if i == j}
        upper = {(r, c - n_users) for r, c in zip(got.row, got.col) if r < n_users}
        lower = {(c, r - n_users) for r, c in zip(got.row, got.col) if r >= n_users}
        assert upper == expect == lower                       # symmetric, and exactly the two-hop pairs
        assert (np.diff(got.row) >= 0).all()                   # row-major entry order


@settings(max_examples=40, deadline=None)
@given(kg_inputs())
def test_kg_adjacency_pair_matches_the_oracle(inp):
    n_users, n_items, n_props, ratings, triples, sym = inp
    ui, ip = build_adjacency_matrix(ratings, np.arange(n_users), np.arange(n_items), props_triples=triples,
                                    props=np.arange(n_props), type_adjacency='unary-kg', symmetric_adjacency=sym)
    oui, oip = og.build_kg_adjacencies(ratings, n_users, n_items, triples, n_props, symmetric=sym)
    for a, b in ((ui, oui), (ip, oip)):
        assert a.shape == b.shape and a.dtype == b.dtype == np.float32
        assert np.array_equal(a.row, b.row) and np.array_equal(a.col, b.col) and np.array_equal(a.data, b.data)
    # the uip graph is the two stacked (the item-property block shifted by the user count): same edge multiset
    uip = build_adjacency_matrix(ratings, np.arange(n_users), np.arange(n_items), props_triples=triples,
                                 props=np.arange(n_props), type_adjacency='unary-uip', symmetric_adjacency=sym)
    stacked = sparse.coo_matrix((np.concatenate([ui.data, ip.data]), (np.concatenate([ui.row, ip.row + n_users]),
                                                                    np.concatenate([ui.col, ip.col + n_users]))), shape=uip.shape)
    assert (abs(uip.tocsr() - stacked.tocsr())).nnz == 0


grid_values = st.lists(st.one_of(st.integers(-3, 3), st.floats(1e-5, 1.0), st.sampled_from(["a", "b", [8, 8], [16, 16, 16]])),
                       min_size=1, max_size=3)


@settings(max_examples=60, deadline=None)
@given(st.dictionaries(st.sampled_from(["model", "dataset", "parameters"]),
                       st.dictionaries(st.sampled_from(["name", "l2_regularizer", "n_hiddens", "epochs", "k"]), grid_values,
                                       min_size=1, max_size=3), min_size=1, max_size=3))
def test_make_grid_is_the_cartesian_product(grid):
    exps = make_grid(grid)
    leaves = [(sec, key, vals) for sec, d in grid.items() for key, vals in d.items()]
    assert len(exps) == int(np.prod([len(v) for _, _, v in leaves]))
    want = [tuple(c) for c in itertools.product(*[v for _, _, v in leaves])]
    got = [tuple(e[sec][key] for sec, key, _ in leaves) for e in exps]
    assert got == want                                         # same order: last listed key varies fastest
    base = {"model": {"name": "x", "other": 1}, "seed": 42}
    for e in exps[:4]:
        merged = nested_dict_update({k: (dict(v) if isinstance(v, dict) else v) for k, v in base.items()}, e)
        assert merged["seed"] == 42 and all(merged[sec][key] == e[sec][key] for sec in e for key in e[sec])
        if "model" in e and "other" not in e["model"]:
            assert merged["model"]["other"] == 1
