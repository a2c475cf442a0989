"""Golden fixtures for the pre-computed-embedding loaders (basic.BasicRS / hybrid.HybridCBRS baselines), from the
REFERENCE's own loaders.  Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_kge.py

data.loaders.load_graph_embeddings / load_bert_embeddings / load_hybrid_embeddings are imported UNMODIFIED from
/root/reference/src (oracle/tf_stub on sys.path, as in make_golden.py) and run on the committed hybrid_small files plus
a small knowledge-graph embedding file written here in the reference's JSON schema
({'ent_embeddings': rows indexed by the original entity id}, loaders.py:85-120).
Output (committed): tests/golden/hybrid_small/{768TransH.json,golden_kge.npz}.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle", "tf_stub"))
sys.path.insert(0, "/root/reference/src")

from data import loaders  # noqa: E402  (the reference's module)


def main():
    root = os.path.join(HERE, "hybrid_small")
    train = np.loadtxt(os.path.join(root, "train2id.tsv"), dtype=np.int64, delimiter="\t")
    n_ent = int(train[:, :2].max()) + 1
    rows = np.round(np.random.RandomState(11).standard_normal((n_ent, 6)) * 0.3, 4)
    with open(os.path.join(root, "768TransH.json"), "w") as fp:
        json.dump({"ent_embeddings": rows.tolist()}, fp)
    base = dict(train_ratings_filepath=os.path.join(root, "train2id.tsv"), test_ratings_filepath=os.path.join(root, "test2id.tsv"),
                train_batch_size=128, test_batch_size=64)
    bert = dict(bert_user_filepath=os.path.join(root, "user-lastlayer.json"), bert_item_filepath=os.path.join(root, "item-lastlayer.json"))
    graph = dict(graph_filepath=os.path.join(root, "768TransH.json"))
    g = {}
    for name, fn, kw in (("graph", loaders.load_graph_embeddings, graph), ("bert", loaders.load_bert_embeddings, bert),
                         ("hybrid", loaders.load_hybrid_embeddings, dict(graph, **bert))):
        tr, te = fn(**base, **kw)
        g[name + "_n_batches"] = np.array([len(tr), len(te)])
        for tag, seq, epochs in (("train", tr, 2), ("test", te, 1)):
            for ep in range(epochs):
                for b in range(len(seq)):
                    x, y = seq[b]
                    for k, arr in enumerate(x):
                        g["%s_%s_ep%d_b%d_x%d" % (name, tag, ep, b, k)] = np.asarray(arr)
                    g["%s_%s_ep%d_b%d_y" % (name, tag, ep, b)] = np.asarray(y)
                seq.on_epoch_end()
        print(name, len(tr), len(te), [np.asarray(a).shape for a in x])
    np.savez_compressed(os.path.join(root, "golden_kge.npz"), **g)


if __name__ == "__main__":
    main()
