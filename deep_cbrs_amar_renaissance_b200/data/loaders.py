"""TSV/JSON loaders (mirror of /root/reference/src/data/loaders.py:11-82,123-144,274-331,383-442).

Same function names, keyword arguments and return values as the reference for
the loaders on the hot path, so `Experimenter.build_dataset`
(experiment.py:120-128) can call them by name with signature-filtered kwargs.
"""
import json

import numpy as np
import pandas as pd

from .datasets import HybridUserItemEmbeddings, UserItemEmbeddings, UserItemGraph, UserItemGraphEmbeddings
from .preprocess import RelationalAdjacency, build_adjacency_matrix, edge_relations, get_user_properties


def _read(path, sep):
    return pd.read_csv(path, sep=sep, header=None).to_numpy()


DEVICE_IDS_MIN_ROWS = 1 << 16  # below this the host numpy path is faster than two H2D/D2H copies


def _on_device(n, device_ids):
    if device_ids is None:
        import torch
        return torch.cuda.is_available() and n >= DEVICE_IDS_MIN_ROWS
    return bool(device_ids)


def _unique_inverse(raw, device_ids=None):
    """np.unique(raw, return_inverse=True) (loaders.py:47-49); on device for large integer columns
    (cbrs_compact_ids: radix sort + scan, bit-identical result)."""
    raw = np.asarray(raw)
    if _on_device(len(raw), device_ids) and np.issubdtype(raw.dtype, np.integer):
        import torch
        from .. import ops
        ids = torch.from_numpy(np.ascontiguousarray(raw, dtype=np.int64)).cuda()
        uniq, inv = ops.compact_ids(ids)
        return uniq.cpu().numpy().astype(raw.dtype), inv.cpu().numpy()
    return np.unique(raw, return_inverse=True)


def _lookup(vocab, raw, what, device_ids=None):
    """Index of each raw id in a sorted-unique vocabulary.

    The reference compares an [n,1] column against the whole vocabulary
    (loaders.py:53-54,64: an n x U boolean, 1.1 GB at MovieLens-1M) and keeps the
    matching column index; a binary search returns the same index in O(n log U).
    """
    raw = np.asarray(raw)
    if _on_device(len(raw), device_ids) and np.issubdtype(raw.dtype, np.integer) and np.issubdtype(np.asarray(vocab).dtype, np.integer):
        import torch
        from .. import ops
        pos = ops.lookup_ids(torch.from_numpy(np.ascontiguousarray(vocab, dtype=np.int64)).cuda(),
                             torch.from_numpy(np.ascontiguousarray(raw, dtype=np.int64)).cuda()).cpu().numpy()
        if (pos < 0).any():
            raise KeyError("{} ids absent from the training vocabulary: {}".format(what, np.unique(raw[pos < 0])[:5]))
        return pos
    pos = np.searchsorted(vocab, raw)
    ok = pos < len(vocab)
    ok[ok] = vocab[pos[ok]] == raw[ok]
    if not ok.all():
        raise KeyError("{} ids absent from the training vocabulary: {}".format(what, np.unique(raw[~ok])[:5]))
    return pos


def load_train_test_ratings(train_filepath, test_filepath, props_filepath=None, sep='\t',
                            return_adjacency=False, type_adjacency='unary', sparse_adjacency=True,
                            symmetric_adjacency=True, device_ids=None, relations=None):
    """Ratings with sequential ids (items offset by the user count) [+ adjacency].
    device_ids: None = compaction on the GPU for columns of >= 65,536 rows when one is present,
    True / False to force; the result is the same array either way."""
    raw_train, raw_test = _read(train_filepath, sep), _read(test_filepath, sep)
    users, u_of = _unique_inverse(raw_train[:, 0], device_ids)
    items, i_of = _unique_inverse(raw_train[:, 1], device_ids)
    n_users = len(users)
    train = np.stack([u_of, i_of + n_users, raw_train[:, 2]], axis=1)
    test = np.stack([_lookup(users, raw_test[:, 0], 'test user', device_ids),
                     _lookup(items, raw_test[:, 1], 'test item', device_ids) + n_users, raw_test[:, 2]], axis=1)
    if not return_adjacency:
        return (train, test), (users, items)

    props = triples = predicates = None
    if type_adjacency in ('unary-kg', 'unary-uip') and props_filepath is not None:
        raw = _read(props_filepath, sep)
        props, p_of = _unique_inverse(raw[:, 1], device_ids)
        # the relation column is dropped, every link weighs one (loaders.py:67-68) ...
        triples = np.stack([_lookup(items, raw[:, 0], 'property item', device_ids), p_of + len(items),
                            np.ones(len(raw), dtype=raw.dtype)], axis=1)
        predicates = raw[:, 2]   # ... unless relations= asks for typed edges (row R, extension; default None = dropped)
    adj = build_adjacency_matrix(train, users, items, props_triples=triples, props=props,
                                 type_adjacency=type_adjacency, sparse_adjacency=sparse_adjacency,
                                 symmetric_adjacency=symmetric_adjacency)
    if relations is not None:
        if type_adjacency != 'unary-uip' or predicates is None or not sparse_adjacency:
            raise ValueError("relations= needs the sparse 'unary-uip' graph and a properties file")
        rel, n_rel = edge_relations(int((train[:, 2] == 1).sum()), predicates, relations, symmetric_adjacency)
        adj = RelationalAdjacency(adj, rel, n_rel)
    return (train, test), (users, items), adj


def load_bert_user_item_embeddings(user_filepath, item_filepath, users, items):
    """[U+I, dim] float32 content table ordered like the node ids (loaders.py:123-144)."""
    def table(path, column, ids):
        df = pd.read_json(path)
        lut = dict(zip(df['ID_OpenKE'].tolist(), df[column].tolist()))
        return np.stack([np.asarray(lut[int(i)], dtype=np.float32) for i in ids])
    return np.concatenate([table(user_filepath, 'profile_embedding', users),
                           table(item_filepath, 'embedding', items)], axis=0)


def load_graph_user_item_embeddings(filepath, users, items):
    """[U+I, dim] float32 rows of the pre-computed knowledge-graph embeddings, users first (loaders.py:85-120):
    the JSON holds {'ent_embeddings': [[...], ...]} indexed by the ORIGINAL (OpenKE) entity id."""
    with open(filepath) as fp:
        table = np.array(json.load(fp)['ent_embeddings'], dtype=np.float32)
    return np.concatenate([table[users], table[items]], axis=0)


def _ratings_only(train_ratings_filepath, test_ratings_filepath, sep):
    return load_train_test_ratings(train_ratings_filepath, test_ratings_filepath, sep=sep, return_adjacency=False)


def load_graph_embeddings(train_ratings_filepath, test_ratings_filepath, graph_filepath, sep='\t', shuffle=True,
                          train_batch_size=1024, test_batch_size=2048):
    """Sequences of pre-computed graph-embedding rows for basic.BasicRS (loaders.py:147-184)."""
    (train, test), (users, items) = _ratings_only(train_ratings_filepath, test_ratings_filepath, sep)
    emb = load_graph_user_item_embeddings(graph_filepath, users, items)
    return (UserItemEmbeddings(train, users, items, emb, batch_size=train_batch_size, shuffle=shuffle),
            UserItemEmbeddings(test, users, items, emb, batch_size=test_batch_size, shuffle=False))


def load_bert_embeddings(train_ratings_filepath, test_ratings_filepath, bert_user_filepath, bert_item_filepath, sep='\t',
                         shuffle=True, train_batch_size=1024, test_batch_size=2048):
    """Sequences of BERT content rows for basic.BasicRS (loaders.py:187-226)."""
    (train, test), (users, items) = _ratings_only(train_ratings_filepath, test_ratings_filepath, sep)
    emb = load_bert_user_item_embeddings(bert_user_filepath, bert_item_filepath, users, items)
    return (UserItemEmbeddings(train, users, items, emb, batch_size=train_batch_size, shuffle=shuffle),
            UserItemEmbeddings(test, users, items, emb, batch_size=test_batch_size, shuffle=False))


def load_hybrid_embeddings(train_ratings_filepath, test_ratings_filepath, graph_filepath, bert_user_filepath,
                           bert_item_filepath, sep='\t', shuffle=True, train_batch_size=1024, test_batch_size=2048):
    """Sequences of (graph, BERT) rows for hybrid.HybridCBRS (loaders.py:229-271)."""
    (train, test), (users, items) = _ratings_only(train_ratings_filepath, test_ratings_filepath, sep)
    graph = load_graph_user_item_embeddings(graph_filepath, users, items)
    bert = load_bert_user_item_embeddings(bert_user_filepath, bert_item_filepath, users, items)
    return (HybridUserItemEmbeddings(train, users, items, graph, bert, batch_size=train_batch_size, shuffle=shuffle),
            HybridUserItemEmbeddings(test, users, items, graph, bert, batch_size=test_batch_size, shuffle=False))


def load_user_item_graph(train_ratings_filepath, test_ratings_filepath, props_triples_filepath=None,
                         sep='\t', type_adjacency='unary', sparse_adjacency=True,
                         symmetric_adjacency=True, user_properties=False, shuffle=True,
                         train_batch_size=1024, test_batch_size=2048, relations=None):
    """relations (extension, row R): None = the reference's untyped graph; 'node-range' / 'predicate' = typed edges
    for basic.BasicRGCN (preprocess.RelationalAdjacency)."""
    (train, test), (users, items), adj = load_train_test_ratings(
        train_ratings_filepath, test_ratings_filepath, props_triples_filepath, sep=sep,
        return_adjacency=True, type_adjacency=type_adjacency, sparse_adjacency=sparse_adjacency,
        symmetric_adjacency=symmetric_adjacency, relations=relations)
    if user_properties and type_adjacency != 'unary-uip':  # loaders.py:319-322 / :427-430
        ui_adj, ip_adj = adj
        adj = (ui_adj, ip_adj, get_user_properties(ui_adj, ip_adj, len(users), len(items)))
    return (UserItemGraph(train, users, items, adj, batch_size=train_batch_size, shuffle=shuffle),
            UserItemGraph(test, users, items, adj, batch_size=test_batch_size, shuffle=False))


def load_user_item_graph_bert_embeddings(train_ratings_filepath, test_ratings_filepath,
                                         bert_user_filepath, bert_item_filepath,
                                         props_triples_filepath=None, sep='\t', type_adjacency='unary',
                                         sparse_adjacency=True, symmetric_adjacency=True, shuffle=True,
                                         train_batch_size=1024, test_batch_size=2048,
                                         user_properties=None):
    (train, test), (users, items), adj = load_train_test_ratings(
        train_ratings_filepath, test_ratings_filepath, props_triples_filepath, sep=sep,
        return_adjacency=True, type_adjacency=type_adjacency, sparse_adjacency=sparse_adjacency,
        symmetric_adjacency=symmetric_adjacency)
    if user_properties and type_adjacency != 'unary-uip':  # loaders.py:319-322 / :427-430
        ui_adj, ip_adj = adj
        adj = (ui_adj, ip_adj, get_user_properties(ui_adj, ip_adj, len(users), len(items)))
    bert = load_bert_user_item_embeddings(bert_user_filepath, bert_item_filepath, users, items)
    return (UserItemGraphEmbeddings(train, users, items, adj, bert, batch_size=train_batch_size, shuffle=shuffle),
            UserItemGraphEmbeddings(test, users, items, adj, bert, batch_size=test_batch_size, shuffle=False))
