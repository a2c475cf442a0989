"""Generate golden fixtures by running the REFERENCE's own numpy/scipy data code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference's data.loaders / data.preprocess / data.datasets /
utilities.math.symmetrize_matrix are imported UNMODIFIED from
/root/reference/src with oracle/tf_stub on sys.path (TensorFlow itself is not
installable here; those modules only need tf.float32 and keras.utils.Sequence).
Outputs (committed): tests/golden/<case>/{*.tsv,*.json,golden.npz}.
Rows pinned: G0 (id compaction), G1 (COO adjacency incl. unary-uip + symmetrise),
S0 (batch contents / gather indices over two epochs), S3 (host BERT gather).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle", "tf_stub"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, REPO)

from data import loaders  # noqa: E402  (the reference's module)
from deep_cbrs_amar_renaissance_b200.data import synthetic  # noqa: E402

CASES = {
    "ui_small": dict(n_users=60, n_items=40, n_ratings=900, seed=42),
    "uip_small": dict(n_users=50, n_items=36, n_ratings=700, seed=7, n_props=30, n_triples=160),
    "hybrid_small": dict(n_users=40, n_items=30, n_ratings=500, seed=3, bert_dim=8),
}


def batches(seq, n_epochs=2):
    out = {}
    for ep in range(n_epochs):
        for b in range(len(seq)):
            x, y = seq[b]
            out["ep%d_b%d_u" % (ep, b)] = np.asarray(x[0])
            out["ep%d_b%d_i" % (ep, b)] = np.asarray(x[1])
            out["ep%d_b%d_y" % (ep, b)] = np.asarray(y)
            if len(x) == 4:
                out["ep%d_b%d_ub" % (ep, b)] = np.asarray(x[2])
                out["ep%d_b%d_ib" % (ep, b)] = np.asarray(x[3])
        seq.on_epoch_end()
    return out


def main():
    for name, cfg in CASES.items():
        root = os.path.join(HERE, name)
        paths = synthetic.write_dataset(root, **cfg)
        kw = dict(train_ratings_filepath=paths["train_ratings_filepath"],
                  test_ratings_filepath=paths["test_ratings_filepath"],
                  train_batch_size=128, test_batch_size=64)
        if "props_triples_filepath" in paths:
            kw.update(props_triples_filepath=paths["props_triples_filepath"], type_adjacency="unary-uip")
        if "bert_user_filepath" in paths:
            kw.update(bert_user_filepath=paths["bert_user_filepath"],
                      bert_item_filepath=paths["bert_item_filepath"])
            train, test = loaders.load_user_item_graph_bert_embeddings(**kw)
        else:
            train, test = loaders.load_user_item_graph(**kw)
        adj = train.adj_matrix
        g = dict(train_ratings=train.ratings, test_ratings=test.ratings, users=train.users,
                 items=train.items, adj_row=adj.row, adj_col=adj.col, adj_data=adj.data,
                 adj_shape=np.array(adj.shape), n_train_batches=np.array(len(train)),
                 n_test_batches=np.array(len(test)))
        csr = adj.tocsr()
        g.update(csr_indptr=csr.indptr, csr_indices=csr.indices, csr_data=csr.data)
        for k, v in batches(train).items():
            g["train_" + k] = v
        for k, v in batches(test, 1).items():
            g["test_" + k] = v
        np.savez_compressed(os.path.join(root, "golden.npz"), **g)
        print(name, "N", adj.shape[0], "coo nnz", adj.nnz, "csr nnz", csr.nnz,
              "max", csr.data.max(), "batches", len(train), len(test))


if __name__ == "__main__":
    main()
