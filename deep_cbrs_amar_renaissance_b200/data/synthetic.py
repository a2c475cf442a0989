"""Seeded synthetic datasets in the reference's on-disk formats.

The real MovieLens-1M files sit behind DVC/Google-Drive (datasets/movielens.dvc)
and are unreachable, so every test and bench uses data of the same SHAPE:
train2id.tsv / test2id.tsv (user, item, rating in {0,1}), props2id-*.tsv
(item, entity, relation) and the BERT JSON schema (ID_OpenKE, embedding /
profile_embedding) read by /root/reference/src/data/loaders.py:43-68,123-144.
"""
import json
import os

import numpy as np

ML1M = dict(n_users=6040, n_items=3706, n_ratings=1000209)


def _zipf_weights(n, exponent, rng):
    w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), exponent)
    rng.shuffle(w)
    return w / w.sum()


def make_ratings(n_users, n_items, n_ratings, seed=42, pos_frac=0.572, test_frac=0.2,
                 id_stride=3, min_per_user=4):
    """Unique (user, item) pairs with Zipf activity / popularity, split 80/20.

    Raw ids are non-contiguous (index * id_stride + offset) so that id compaction
    is exercised.  Every test user and item also occurs in train
    (datasets/README.md:21-22).  Returns (train [R,3], test [T,3]) int64.
    """
    rng = np.random.RandomState(seed)
    n_ratings = min(n_ratings, n_users * n_items // 2)
    pu = _zipf_weights(n_users, 0.8, rng)
    pi = _zipf_weights(n_items, 0.9, rng)
    # floor: every user rates at least min_per_user items, every item is rated once
    base_u = np.repeat(np.arange(n_users), min_per_user)
    base_i = rng.randint(0, n_items, size=len(base_u))
    cover_i = np.arange(n_items)
    cover_u = rng.randint(0, n_users, size=n_items)
    keys = np.unique(np.concatenate([base_u, cover_u]).astype(np.int64) * n_items
                     + np.concatenate([base_i, cover_i]))
    while len(keys) < n_ratings:
        need = int((n_ratings - len(keys)) * 1.3) + 16
        u = rng.choice(n_users, size=need, p=pu)
        i = rng.choice(n_items, size=need, p=pi)
        keys = np.unique(np.concatenate([keys, u.astype(np.int64) * n_items + i]))
    if len(keys) > n_ratings:
        keys = np.sort(rng.choice(keys, size=n_ratings, replace=False))
    u, i = keys // n_items, keys % n_items
    y = (rng.random_sample(len(keys)) < pos_frac).astype(np.int64)
    perm = rng.permutation(len(keys))
    u, i, y = u[perm], i[perm], y[perm]
    # first occurrence of every user and of every item is forced into train
    first_u = np.zeros(len(u), bool)
    first_u[np.unique(u, return_index=True)[1]] = True
    first_i = np.zeros(len(u), bool)
    first_i[np.unique(i, return_index=True)[1]] = True
    is_test = (rng.random_sample(len(u)) < test_frac) & ~first_u & ~first_i
    raw_u = u * id_stride + 1
    raw_i = i * id_stride + 2 + n_users * id_stride
    data = np.stack([raw_u, raw_i, y], axis=1).astype(np.int64)
    return data[~is_test], data[is_test]


def make_props(train, n_props, n_triples, seed=42, n_relations=11, dup_frac=0.01):
    """(item, entity, relation) triples; >= dup_frac of (item, entity) pairs are
    repeated under a second predicate (duplicates stay duplicate COO
    entries and are summed by gcn_filter / kept by GAT and GraphSage)."""
    rng = np.random.RandomState(seed + 1)
    items = np.unique(train[:, 1])
    pp = _zipf_weights(n_props, 0.7, rng)
    n_base = int(n_triples * (1 - dup_frac))
    it = items[rng.randint(0, len(items), size=n_base)]
    en = rng.choice(n_props, size=n_base, p=pp) * 5 + 10_000_000
    rel = rng.randint(0, n_relations, size=n_base)
    n_dup = n_triples - n_base
    pick = rng.randint(0, n_base, size=n_dup)
    it = np.concatenate([it, it[pick]])
    en = np.concatenate([en, en[pick]])
    rel = np.concatenate([rel, (rel[pick] + 1) % n_relations])
    perm = rng.permutation(len(it))
    return np.stack([it[perm], en[perm], rel[perm]], axis=1).astype(np.int64)


def make_bert(n_rows, dim=768, seed=42, scale=0.5):
    """BERT-like content rows ~ N(0,1)*scale, float32 (summed-word-embedding scale)."""
    rng = np.random.RandomState(seed + 2)
    return (rng.standard_normal((n_rows, dim)) * scale).astype(np.float32)


def write_tsv(path, arr):
    np.savetxt(path, arr, fmt="%d", delimiter="\t")


def write_bert_json(path, ids, rows, column):
    """Schema of embeddings/bert/*.json as read by loaders.py:104-105,135-142."""
    recs = [{"ID_OpenKE": int(i), column: [float(v) for v in r]} for i, r in zip(ids, rows)]
    with open(path, "w") as fp:
        json.dump(recs, fp)


def write_dataset(root, n_users, n_items, n_ratings, seed=42, n_props=0, n_triples=0,
                  bert_dim=0):
    """Write a full synthetic dataset directory; returns the dict of file paths."""
    os.makedirs(root, exist_ok=True)
    train, test = make_ratings(n_users, n_items, n_ratings, seed)
    paths = dict(train_ratings_filepath=os.path.join(root, "train2id.tsv"),
                 test_ratings_filepath=os.path.join(root, "test2id.tsv"))
    write_tsv(paths["train_ratings_filepath"], train)
    write_tsv(paths["test_ratings_filepath"], test)
    if n_props:
        props = make_props(train, n_props, n_triples, seed)
        paths["props_triples_filepath"] = os.path.join(root, "props2id.tsv")
        write_tsv(paths["props_triples_filepath"], props)
    if bert_dim:
        users, items = np.unique(train[:, 0]), np.unique(train[:, 1])
        paths["bert_user_filepath"] = os.path.join(root, "user-lastlayer.json")
        paths["bert_item_filepath"] = os.path.join(root, "item-lastlayer.json")
        write_bert_json(paths["bert_user_filepath"], users,
                        make_bert(len(users), bert_dim, seed), "profile_embedding")
        write_bert_json(paths["bert_item_filepath"], items,
                        make_bert(len(items), bert_dim, seed + 7), "embedding")
    return paths
