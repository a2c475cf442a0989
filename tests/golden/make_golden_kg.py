"""Golden fixtures for the Two-Step / Two-Way graphs (scope row (f)-4), from the REFERENCE's own loader.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_kg.py

data.loaders.load_user_item_graph is imported UNMODIFIED from /root/reference/src (oracle/tf_stub on sys.path, as in
make_golden.py) and run on the committed uip_small TSV files with type_adjacency='unary-kg', user_properties=True:
`trainset.adj_matrix` is then (user-item [U+I], item-property [I+P], user-property [U+P]) - the last one built by the
reference's get_user_properties (preprocess.py:9-41) through its dense [U+I+P]^2 detour.
Output (committed): tests/golden/uip_small/golden_kg.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle", "tf_stub"))
sys.path.insert(0, "/root/reference/src")

from data import loaders  # noqa: E402  (the reference's module)


def main():
    root = os.path.join(HERE, "uip_small")
    kw = dict(train_ratings_filepath=os.path.join(root, "train2id.tsv"), test_ratings_filepath=os.path.join(root, "test2id.tsv"),
              props_triples_filepath=os.path.join(root, "props2id.tsv"), type_adjacency="unary-kg",
              train_batch_size=128, test_batch_size=64)
    g = {}
    for sym in (True, False):
        train, _ = loaders.load_user_item_graph(user_properties=True, symmetric_adjacency=sym, **kw)
        tag = "sym" if sym else "dir"
        for name, m in zip(("ui", "ip", "up"), train.adj_matrix):
            g["%s_%s_row" % (tag, name)], g["%s_%s_col" % (tag, name)], g["%s_%s_data" % (tag, name)] = m.row, m.col, m.data
            g["%s_%s_shape" % (tag, name)] = np.array(m.shape)
            print(tag, name, m.shape, m.nnz, m.dtype, m.row.dtype)
    g["n_users"], g["n_items"] = np.array(len(train.users)), np.array(len(train.items))
    np.savez_compressed(os.path.join(root, "golden_kg.npz"), **g)


if __name__ == "__main__":
    main()
