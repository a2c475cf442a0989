"""Adjacency construction on the host (mirror of /root/reference/src/data/preprocess.py:44-170).

Only the adjacency types on the hot path are built: 'unary' and 'unary-uip'
(SURVEY.md section 2 row 1); 'binary' is kept because it is one line; 'unary-kg' and
get_user_properties belong to the Two-Step / Two-Way variants and are out of
scope (they raise).
"""
import numpy as np
from scipy import sparse

from ..utilities.math import symmetrize_matrix


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
        raise NotImplementedError("'unary-kg' feeds only the Two-Step/Two-Way variants (out of scope, DESIGN.md)")
    raise ValueError("Unknown adjacency matrix type named {}".format(type_adjacency))
