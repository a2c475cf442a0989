"""Adjacency construction on the host (mirror of /root/reference/src/data/preprocess.py:44-170).

Only the adjacency types on the hot path are built: 'unary' and 'unary-uip'
(SURVEY.md section 2 row 1); 'binary' is kept because it is one line; 'unary-kg' (two graphs)
feeds the Two-Step variants, get_user_properties (the third graph) the Two-Way variants (scope row (f)-4).
"""
import numpy as np
from scipy import sparse

from ..utilities.math import symmetrize_matrix


def get_user_properties(ui_adj, ip_adj, n_users, n_items):
    """User-property adjacency over [U+P] nodes: user u and property p are linked when some item is linked to both
    (preprocess.py:9-41).  The reference squares the stacked [U+I+P] adjacency, sets every stored value to one,
    DENSIFIES the square and copies two blocks into a dense [U+P, U+P] array - the reason the Two-Way variants run out
    of memory on real graphs.  Here everything stays sparse; the result is the same COO matrix: float64 ones, entries
    in row-major order (what sparse.coo_matrix(dense) yields)."""
    n_props = ip_adj.shape[0] - n_items
    n = n_users + n_items + n_props
    ui, ip = ui_adj.tocoo(), ip_adj.tocoo()
    uip = sparse.coo_matrix((np.concatenate([ui.data, ip.data]),
                             (np.concatenate([ui.row, ip.row + n_users]), np.concatenate([ui.col, ip.col + n_users]))),
                            shape=(n, n)).tocsr()
    two_hop = uip.dot(uip).tocsr()
    two_hop.eliminate_zeros()  # a stored zero is a zero in the reference's dense copy
    two_hop.data = np.ones(len(two_hop.data))
    up = sparse.bmat([[None, two_hop[:n_users, n_users + n_items:]],
                      [two_hop[n_users + n_items:, :n_users], None]], format='csr', dtype=np.float64)
    up.sort_indices()
    return up.tocoo()


class RelationalAdjacency:
    """A COO adjacency plus one relation id per stored entry (scope row R, extension).

    The reference reads the predicate column of props2id-*.tsv and drops it (/root/reference/src/data/loaders.py:63-68):
    its user-item-properties graphs are untyped.  `load_user_item_graph(..., relations=...)` keeps a type per edge
    instead and hands the models this wrapper; everything that takes the reference's scipy matrix still works on
    `.coo` (and `tocoo()` / `.shape`), so the non-relational layer families run on the same object unchanged.
      relations='node-range': 0 = user-item edge, 1 = item-property edge (defined by the node-id ranges of
                              /root/reference/src/data/preprocess.py:149-159);
      relations='predicate' : 0 = user-item edge, 1 + p = item-property edge with the p-th distinct predicate.
    Both directions of a symmetrised edge carry the same relation, so every relation block is symmetric."""

    def __init__(self, coo, rel, n_rel):
        self.coo, self.rel, self.n_rel = coo, np.ascontiguousarray(rel, dtype=np.int32), int(n_rel)
        if len(self.rel) != coo.nnz:
            raise ValueError("one relation id per stored entry: {} ids for {} entries".format(len(self.rel), coo.nnz))
        self.shape = coo.shape

    def tocoo(self):
        return self.coo

    def relation_blocks(self):
        """[scipy COO per relation] (tests, oracle)"""
        c = self.coo
        return [sparse.coo_matrix((c.data[self.rel == r], (c.row[self.rel == r], c.col[self.rel == r])), shape=c.shape)
                for r in range(self.n_rel)]


def edge_relations(n_liked, props_triples_rel, relations, symmetric):
    """relation id of every entry of the 'unary-uip' COO in its entry order: rating edges, property edges, then (when
    symmetric) the transposed entries in the same order.  props_triples_rel: the predicate column of the triples."""
    if relations == 'node-range':
        rel_p = np.ones(len(props_triples_rel), dtype=np.int32)
        n_rel = 2
    elif relations == 'predicate':
        preds, inv = np.unique(props_triples_rel, return_inverse=True)
        rel_p = (1 + inv).astype(np.int32)
        n_rel = 1 + len(preds)
    else:
        raise ValueError("relations must be None, 'node-range' or 'predicate', got {!r}".format(relations))
    one = np.concatenate([np.zeros(n_liked, dtype=np.int32), rel_p])
    return (np.concatenate([one, one]) if symmetric else one), n_rel


def _coo(data, rows, cols, n, symmetric, as_sparse):
    m = sparse.coo_matrix((data, (rows, cols)), shape=[n, n], dtype=np.float32)
    if symmetric:
        m = symmetrize_matrix(m)
    return m if as_sparse else m.todense()


def build_adjacency_matrix(bi_ratings, users, items, props_triples=None, props=None,
                           type_adjacency='unary', sparse_adjacency=True, symmetric_adjacency=True):
    """Same arguments and results as preprocess.py:44-170 for 'unary', 'binary', 'unary-uip'.

    Node numbering: users [0,U), items [U,U+I), properties [U+I,U+I+P).  The
    ratings triples already carry item ids offset by U; property triples carry
    (item index, len(items)+prop index) and are shifted by U here
    (preprocess.py:150-151).  Entry order: rating edges, then property edges,
    then (when symmetric) all transposed entries.
    """
    n_ui = len(users) + len(items)
    if type_adjacency == 'binary':
        return _coo(bi_ratings[:, 2], bi_ratings[:, 0], bi_ratings[:, 1], n_ui,
                    symmetric_adjacency, sparse_adjacency)
    liked = bi_ratings[bi_ratings[:, 2] == 1]
    if type_adjacency == 'unary':
        return _coo(liked[:, 2], liked[:, 0], liked[:, 1], n_ui, symmetric_adjacency, sparse_adjacency)
    if type_adjacency == 'unary-uip':
        if props is None or props_triples is None:
            raise ValueError("KG adjacency matrix requires properties info")
        shift = len(users)
        return _coo(np.concatenate([liked[:, 2], props_triples[:, 2]]),
                    np.concatenate([liked[:, 0], props_triples[:, 0] + shift]),
                    np.concatenate([liked[:, 1], props_triples[:, 1] + shift]),
                    n_ui + len(props), symmetric_adjacency, sparse_adjacency)
    if type_adjacency == 'unary-kg':
        # two graphs (preprocess.py:113-145): user-item over [U+I] and item-property over [I+P] (items first)
        if props is None or props_triples is None:
            raise ValueError("KG adjacency matrix requires properties info")
        return (_coo(liked[:, 2], liked[:, 0], liked[:, 1], n_ui, symmetric_adjacency, sparse_adjacency),
                _coo(props_triples[:, 2], props_triples[:, 0], props_triples[:, 1], len(items) + len(props),
                     symmetric_adjacency, sparse_adjacency))
    raise ValueError("Unknown adjacency matrix type named {}".format(type_adjacency))
