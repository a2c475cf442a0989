"""TSV/JSON loaders (mirror of /root/reference/src/data/loaders.py:11-82,123-144,274-331,383-442).

Same function names, keyword arguments and return values as the reference for
the loaders on the hot path, so `Experimenter.build_dataset`
(experiment.py:120-128) can call them by name with signature-filtered kwargs.
"""
import numpy as np
import pandas as pd

from .datasets import UserItemGraph, UserItemGraphEmbeddings
from .preprocess import build_adjacency_matrix


def _read(path, sep):
    return pd.read_csv(path, sep=sep, header=None).to_numpy()


def _lookup(vocab, raw, what):
    """Index of each raw id in a sorted-unique vocabulary.

    The reference compares an [n,1] column against the whole vocabulary
    (loaders.py:53-54,64: an n x U boolean, 1.1 GB at MovieLens-1M) and keeps the
    matching column index; a binary search returns the same index in O(n log U).
    """
    pos = np.searchsorted(vocab, raw)
    ok = pos < len(vocab)
    ok[ok] = vocab[pos[ok]] == raw[ok]
    if not ok.all():
        raise KeyError("{} ids absent from the training vocabulary: {}".format(what, np.unique(raw[~ok])[:5]))
    return pos


def load_train_test_ratings(train_filepath, test_filepath, props_filepath=None, sep='\t',
                            return_adjacency=False, type_adjacency='unary', sparse_adjacency=True,
                            symmetric_adjacency=True):
    """Ratings with sequential ids (items offset by the user count) [+ adjacency]."""
    raw_train, raw_test = _read(train_filepath, sep), _read(test_filepath, sep)
    users, u_of = np.unique(raw_train[:, 0], return_inverse=True)
    items, i_of = np.unique(raw_train[:, 1], return_inverse=True)
    n_users = len(users)
    train = np.stack([u_of, i_of + n_users, raw_train[:, 2]], axis=1)
    test = np.stack([_lookup(users, raw_test[:, 0], 'test user'),
                     _lookup(items, raw_test[:, 1], 'test item') + n_users, raw_test[:, 2]], axis=1)
    if not return_adjacency:
        return (train, test), (users, items)

    props = triples = None
    if type_adjacency in ('unary-kg', 'unary-uip') and props_filepath is not None:
        raw = _read(props_filepath, sep)
        props, p_of = np.unique(raw[:, 1], return_inverse=True)
        # the relation column is dropped, every link weighs one (loaders.py:67-68)
        triples = np.stack([_lookup(items, raw[:, 0], 'property item'), p_of + len(items),
                            np.ones(len(raw), dtype=raw.dtype)], axis=1)
    adj = build_adjacency_matrix(train, users, items, props_triples=triples, props=props,
                                 type_adjacency=type_adjacency, sparse_adjacency=sparse_adjacency,
                                 symmetric_adjacency=symmetric_adjacency)
    return (train, test), (users, items), adj


def load_bert_user_item_embeddings(user_filepath, item_filepath, users, items):
    """[U+I, dim] float32 content table ordered like the node ids (loaders.py:123-144)."""
    def table(path, column, ids):
        df = pd.read_json(path)
        lut = dict(zip(df['ID_OpenKE'].tolist(), df[column].tolist()))
        return np.stack([np.asarray(lut[int(i)], dtype=np.float32) for i in ids])
    return np.concatenate([table(user_filepath, 'profile_embedding', users),
                           table(item_filepath, 'embedding', items)], axis=0)


def load_user_item_graph(train_ratings_filepath, test_ratings_filepath, props_triples_filepath=None,
                         sep='\t', type_adjacency='unary', sparse_adjacency=True,
                         symmetric_adjacency=True, user_properties=False, shuffle=True,
                         train_batch_size=1024, test_batch_size=2048):
    if user_properties and type_adjacency != 'unary-uip':
        raise NotImplementedError("user-properties graphs feed only the Two-Way variant (out of scope)")
    (train, test), (users, items), adj = load_train_test_ratings(
        train_ratings_filepath, test_ratings_filepath, props_triples_filepath, sep=sep,
        return_adjacency=True, type_adjacency=type_adjacency, sparse_adjacency=sparse_adjacency,
        symmetric_adjacency=symmetric_adjacency)
    return (UserItemGraph(train, users, items, adj, batch_size=train_batch_size, shuffle=shuffle),
            UserItemGraph(test, users, items, adj, batch_size=test_batch_size, shuffle=False))


def load_user_item_graph_bert_embeddings(train_ratings_filepath, test_ratings_filepath,
                                         bert_user_filepath, bert_item_filepath,
                                         props_triples_filepath=None, sep='\t', type_adjacency='unary',
                                         sparse_adjacency=True, symmetric_adjacency=True, shuffle=True,
                                         train_batch_size=1024, test_batch_size=2048,
                                         user_properties=None):
    if user_properties and type_adjacency != 'unary-uip':
        raise NotImplementedError("user-properties graphs feed only the Two-Way variant (out of scope)")
    (train, test), (users, items), adj = load_train_test_ratings(
        train_ratings_filepath, test_ratings_filepath, props_triples_filepath, sep=sep,
        return_adjacency=True, type_adjacency=type_adjacency, sparse_adjacency=sparse_adjacency,
        symmetric_adjacency=symmetric_adjacency)
    bert = load_bert_user_item_embeddings(bert_user_filepath, bert_item_filepath, users, items)
    return (UserItemGraphEmbeddings(train, users, items, adj, bert, batch_size=train_batch_size, shuffle=shuffle),
            UserItemGraphEmbeddings(test, users, items, adj, bert, batch_size=test_batch_size, shuffle=False))
