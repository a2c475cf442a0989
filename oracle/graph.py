"""Oracle (TEST INFRASTRUCTURE): graph build, rows G0-G3 of SURVEY.md section 8a.

Every function cites the reference lines it restates.  Structure outputs
(indptr / indices / id maps) are integer and must match the product bit-for-bit.
"""
import numpy as np
from scipy import sparse


# --------------------------------------------------------------------------- G0
def compact_ids(train_ratings, test_ratings):
    """Id compaction of the (user, item, rating) triples.

    Restates /root/reference/src/data/loaders.py:48-56:
      users, u_idx = np.unique(train[:,0], return_inverse=True); same for items;
      item indexes are offset by len(users); test ids are mapped through the
      train vocabularies (the reference does it with an O(n_test*U) broadcast +
      argwhere, loaders.py:53-54; searchsorted gives the same indices because
      the vocabularies are sorted-unique).
    A test id absent from train makes the reference's np.stack fail on ragged
    lengths; here it raises KeyError.
    """
    train_ratings = np.asarray(train_ratings)
    test_ratings = np.asarray(test_ratings)
    users, u_idx = np.unique(train_ratings[:, 0], return_inverse=True)
    items, i_idx = np.unique(train_ratings[:, 1], return_inverse=True)
    i_idx = i_idx + len(users)
    train = np.stack([u_idx, i_idx, train_ratings[:, 2]], axis=1)

    tu = np.searchsorted(users, test_ratings[:, 0])
    ti = np.searchsorted(items, test_ratings[:, 1])
    if (tu >= len(users)).any() or (users[np.minimum(tu, len(users) - 1)] != test_ratings[:, 0]).any():
        raise KeyError("test user id absent from train")
    if (ti >= len(items)).any() or (items[np.minimum(ti, len(items) - 1)] != test_ratings[:, 1]).any():
        raise KeyError("test item id absent from train")
    test = np.stack([tu, ti + len(users), test_ratings[:, 2]], axis=1)
    return train, test, users, items


def compact_props(props_triples, items):
    """Property-triple mapping; the relation column is replaced by ones.

    Restates loaders.py:62-68: item ids -> item index (no user offset yet),
    props = np.unique(col1), prop index += len(items), data := 1.
    Returns (triples [T,3] int64, props).
    """
    props_triples = np.asarray(props_triples)
    ii = np.searchsorted(items, props_triples[:, 0])
    if (ii >= len(items)).any() or (items[np.minimum(ii, len(items) - 1)] != props_triples[:, 0]).any():
        raise KeyError("property item id absent from train items")
    props, p_idx = np.unique(props_triples[:, 1], return_inverse=True)
    p_idx = p_idx + len(items)
    ones = np.ones(len(p_idx), dtype=props_triples.dtype)
    return np.stack([ii, p_idx, ones], axis=1), props


# --------------------------------------------------------------------------- G1
def build_adjacency(train, n_users, n_items, props_triples=None, n_props=0,
                    type_adjacency="unary", symmetric=True):
    """COO adjacency (row, col int32; data float32) in the reference's entry order.

    Restates preprocess.py:68-86 ('unary') and :113-127,149-163 ('unary-uip'),
    plus symmetrize_matrix utilities/math.py:13-20: positives only, value = the
    rating (1), property edges appended after rating edges with +n_users offset,
    then (row||col, col||row) concatenation WITHOUT dedup.
    """
    train = np.asarray(train)
    pos = train[:, 2] == 1
    rows, cols, data = train[pos, 0], train[pos, 1], train[pos, 2]
    n = n_users + n_items
    if type_adjacency == "unary-uip":
        if props_triples is None:
            raise ValueError("KG adjacency matrix requires properties info")
        rows = np.concatenate([rows, props_triples[:, 0] + n_users])
        cols = np.concatenate([cols, props_triples[:, 1] + n_users])
        data = np.concatenate([data, props_triples[:, 2]])
        n += n_props
    elif type_adjacency != "unary":
        raise ValueError("Unknown adjacency matrix type named {}".format(type_adjacency))
    rows = rows.astype(np.int32)
    cols = cols.astype(np.int32)
    data = data.astype(np.float32)
    if symmetric:
        rows, cols = np.concatenate([rows, cols]), np.concatenate([cols, rows])
        data = np.concatenate([data, data])
    return sparse.coo_matrix((data, (rows, cols)), shape=(n, n), dtype=np.float32)


def build_kg_adjacencies(train, n_users, n_items, props_triples, n_props, symmetric=True):
    """'unary-kg' (preprocess.py:113-145): the user-item graph over [U+I] and the item-property graph over [I+P]
    (items first: triples carry (item index, n_items + property index)), each symmetrised separately."""
    ui = build_adjacency(train, n_users, n_items, symmetric=symmetric)
    rows, cols = props_triples[:, 0].astype(np.int32), props_triples[:, 1].astype(np.int32)
    data = props_triples[:, 2].astype(np.float32)
    if symmetric:
        rows, cols = np.concatenate([rows, cols]), np.concatenate([cols, rows])
        data = np.concatenate([data, data])
    n = n_items + n_props
    return ui, sparse.coo_matrix((data, (rows, cols)), shape=(n, n), dtype=np.float32)


def user_properties(ui_adj, ip_adj, n_users, n_items):
    """get_user_properties (preprocess.py:9-41), restated with its dense detour (small graphs only): stack the two
    graphs over [U+I+P], square, set every stored value to one, densify, copy the property x user and the
    user x property blocks into a dense [U+P]^2 array, and return its COO form (float64, row-major order)."""
    n_props = ip_adj.shape[0] - n_items
    n = n_users + n_items + n_props
    stacked = sparse.coo_matrix((np.concatenate([ui_adj.data, ip_adj.data]),
                                 (np.concatenate([ui_adj.row, ip_adj.row + n_users]),
                                  np.concatenate([ui_adj.col, ip_adj.col + n_users]))), shape=(n, n))
    sq = stacked.dot(stacked)
    sq.data = np.ones(len(sq.data))
    sq = np.asarray(sq.todense())
    up = np.zeros((n_users + n_props, n_users + n_props))
    up[n_users:, :n_users] = sq[n_users + n_items:, :n_users]
    up[:n_users, n_users:] = sq[:n_users, n_users + n_items:]
    return sparse.coo_matrix(up)


# --------------------------------------------------------------------------- G2
def inv_sqrt_degree(deg):
    """d = deg^(-1/2) as float32, inf -> 0.

    [3P] spektral.utils.degree_power: np.power(A.sum(1), -0.5) on float32.
    Defined here as the correctly rounded float32 of the float64 value; glibc's
    powf (what numpy calls) agrees with this on every integer degree tested
    (tests/test_oracle_graph.py::test_inv_sqrt_matches_numpy_power).
    """
    deg = np.asarray(deg, dtype=np.float32)
    with np.errstate(divide="ignore"):
        d = (1.0 / np.sqrt(deg.astype(np.float64))).astype(np.float32)
    d[np.isinf(d)] = 0.0
    return d


def gcn_filter(adj_coo):
    """Normalised adjacency with self loops, CSR float32, ascending columns.

    [3P] spektral.utils.gcn_filter(A, symmetric=True), call sites
    /root/reference/src/models/gnn.py:283 and src/layers/lightgcn_conv.py:58:
      M = A.tocsr() (duplicates summed); M_ii += 1; deg = row sums;
      d = deg^-1/2; A_hat_ij = fl32(fl32(d_i * M_ij) * d_j); sort_indices().
    """
    m = sparse.csr_matrix(adj_coo, dtype=np.float32)
    m.sum_duplicates()
    n = m.shape[0]
    m = (m + sparse.identity(n, dtype=np.float32, format="csr")).tocsr()
    m.sort_indices()
    deg = np.asarray(m.sum(axis=1), dtype=np.float32).ravel()
    d = inv_sqrt_degree(deg)
    row_of = np.repeat(np.arange(n), np.diff(m.indptr))
    left = (d[row_of] * m.data).astype(np.float32)
    vals = (left * d[m.indices]).astype(np.float32)
    out = sparse.csr_matrix((vals, m.indices.copy(), m.indptr.copy()), shape=m.shape)
    return out


def gcn_filter_scipy(adj_coo):
    """The same filter written the way Spektral writes it (scipy diags products).

    Used only to cross-check `gcn_filter` (values may differ by 1 ulp where
    numpy's float32 power is not correctly rounded).
    """
    out = adj_coo.tocsr().astype(np.float32)
    out = out.tolil()
    out.setdiag(out.diagonal() + 1)
    out = out.tocsr()
    deg = np.power(np.array(out.sum(1), dtype=np.float32), np.float32(-0.5)).ravel()
    deg[np.isinf(deg)] = 0.0
    dm = sparse.diags(deg.astype(np.float32))
    out = dm.dot(out).dot(dm).tocsr()
    out.sort_indices()
    return out


# --------------------------------------------------------------------------- G3
def reorder_raw(adj_coo):
    """Row-major sorted edge list WITH duplicates kept, as CSR arrays.

    Restates utilities/math.py:47-56 (tocoo -> SparseTensor -> tf.sparse.reorder):
    what GAT and GraphSage layers receive (gnn.py:298-319,331-352 do not
    preprocess).  Stable sort by (row, col).
    Returns (indptr int64 [N+1], indices int32 [nnz], data float32 [nnz]).
    """
    coo = adj_coo.tocoo()
    order = np.lexsort((coo.col, coo.row))  # lexsort is stable
    rows = coo.row[order]
    n = coo.shape[0]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, rows.astype(np.int64) + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, coo.col[order].astype(np.int32), coo.data[order].astype(np.float32)


def gat_edges(indptr, indices):
    """GAT edge set: drop existing self loops, add (i,i) for all i, re-sort.

    [3P] spektral.layers.ops.add_self_loops_indices, called by GATConv with
    add_self_loops=True (default; built at gnn.py:321-328).
    """
    n = len(indptr) - 1
    row_of = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    keep = row_of != indices
    rows = np.concatenate([row_of[keep], np.arange(n, dtype=np.int64)])
    cols = np.concatenate([indices[keep].astype(np.int64), np.arange(n, dtype=np.int64)])
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    out_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(out_ptr, rows + 1, 1)
    return np.cumsum(out_ptr), cols.astype(np.int32)
