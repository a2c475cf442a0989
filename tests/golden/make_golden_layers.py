"""Golden vectors from the REFERENCE's own layer / metric code, run unmodified.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_layers.py

src/layers/reduction.py, src/layers/fusion.py, src/layers/dgcf_conv.py are imported UNMODIFIED from /root/reference/src
with oracle/tf_np_stub on sys.path: their arithmetic is the reference's own Python over generic array ops
(tf.concat / add_n / matmul / tanh / softmax / reduce_sum / argmin ...), which the stub maps one-to-one onto numpy
float32 functions.  src/utilities/metrics.py:top_k_predictions is pure pandas; it needs DataFrame.append, removed in
pandas 2, which is restored here as the one-line pd.concat it always was.
Pinned by these vectors: row P5 (all five reductions), FusionLayer('attention') incl. both projection cases, the DGCF
operator recipe and layer (row (f)-3; through the oracle's gcn_filter restatement, see oracle/tf_np_stub/README.md),
row T (pair-list top-k incl. ties).  Output (committed): tests/golden/layers/golden_layers.npz.
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle", "tf_np_stub"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, REPO)

if not hasattr(pd.DataFrame, "append"):  # pandas >= 2: DataFrame.append(other) == pd.concat([self, other])
    pd.DataFrame.append = lambda self, other: pd.concat([self, other])

from layers import dgcf_conv, fusion, reduction  # noqa: E402  (the reference's modules)
from utilities import metrics  # noqa: E402
from tests.helpers import random_bipartite  # noqa: E402


def main():
    rng = np.random.RandomState(2024)
    g = {}
    # ---- reductions (src/layers/reduction.py:5-55) -----------------------------------------------------
    hs = [rng.standard_normal((20, 6)).astype(np.float32) for _ in range(3)]
    for l, h in enumerate(hs):
        g["red_h%d" % l] = h
    for method in ("concatenation", "sum", "mean", "last"):
        g["red_" + method] = np.asarray(reduction.ReductionLayer(method)(hs))
    ws = reduction.ReductionLayer("w-sum")
    ws(hs)                                       # builds the [L,1,1] weights (ones)
    ws.layer.w[...] = rng.uniform(0.5, 1.5, size=ws.layer.w.shape).astype(np.float32)
    g["red_wsum_w"] = ws.layer.w.copy()
    g["red_w-sum"] = np.asarray(ws(hs))
    # ---- attention fusion (src/layers/fusion.py:19-68) ---------------------------------------------------
    for tag, fa, fb in (("same", 8, 8), ("projA", 6, 10), ("projB", 10, 6)):
        a = rng.standard_normal((16, fa)).astype(np.float32)
        b = rng.standard_normal((16, fb)).astype(np.float32)
        layer = fusion.FusionLayer("attention")
        out = layer([a, b])
        g["fus_%s_a" % tag], g["fus_%s_b" % tag], g["fus_%s_out" % tag] = a, b, np.asarray(out)
        g["fus_%s_att" % tag] = layer.att_weight
        if layer.proj_first is not None:
            g["fus_%s_proj" % tag] = layer.proj_weight
            g["fus_%s_proj_first" % tag] = np.array(bool(layer.proj_first))
    a = rng.standard_normal((5, 3)).astype(np.float32)
    b = rng.standard_normal((5, 4)).astype(np.float32)
    g["fus_cat_a"], g["fus_cat_b"], g["fus_cat_out"] = a, b, np.asarray(fusion.FusionLayer("concatenate")([a, b]))
    # ---- DGCF operator + layer (src/layers/dgcf_conv.py:32-80,101-102) ------------------------------------
    for tag, adj in (("ui", random_bipartite(40, 30, 320, seed=5)),
                     ("uip", random_bipartite(30, 24, 260, seed=6, n_props=12, n_links=60, dup_links=15))):
        m = dgcf_conv.DGCFConv.preprocess(adj).tocsr()
        m.sort_indices()
        g["dgcf_%s_row" % tag], g["dgcf_%s_col" % tag], g["dgcf_%s_val" % tag] = adj.row, adj.col, adj.data
        g["dgcf_%s_n" % tag] = np.array(adj.shape[0])
        g["dgcf_%s_indptr" % tag], g["dgcf_%s_indices" % tag], g["dgcf_%s_data" % tag] = m.indptr, m.indices, m.data.astype(np.float32)
        layer = dgcf_conv.DGCFConv(None)
        x = rng.standard_normal((adj.shape[0], 5)).astype(np.float32)
        layer([x, m])                                   # builds the gate (ones)
        layer.locality_adaptive.w[...] = rng.standard_normal(layer.locality_adaptive.w.shape).astype(np.float32)
        g["dgcf_%s_x" % tag], g["dgcf_%s_w" % tag] = x, layer.locality_adaptive.w.copy()
        g["dgcf_%s_out" % tag] = np.asarray(layer([x, m]))
    # ---- pair-list top-k (src/utilities/metrics.py:11-34) -------------------------------------------------------
    users = np.array([101, 205, 309, 412, 777])
    items = np.array([11, 22, 33, 44, 55, 66, 77])
    n = 60
    pu = rng.randint(0, len(users), size=n)
    pi = rng.randint(0, len(items), size=n) + len(users)
    sc = np.round(rng.uniform(0, 1, size=n), 1)          # one decimal: plenty of exact ties
    preds = np.stack([pu, pi, sc], axis=1)
    g["topk_preds"], g["topk_users"], g["topk_items"] = preds, users, items
    for k in (1, 3, 5):
        df = metrics.top_k_predictions(preds, users, items, k=k)
        # the reference walks a Python set of users (arbitrary order); inside a user the order is its own: keep it
        df = pd.concat([df[df["users"] == u] for u in sorted(set(df["users"]))])
        g["topk_k%d_users" % k] = df["users"].to_numpy()
        g["topk_k%d_items" % k] = df["items"].to_numpy()
        g["topk_k%d_scores" % k] = df["scores"].to_numpy()
    np.savez_compressed(os.path.join(HERE, "layers", "golden_layers.npz"), **g)
    print("wrote", len(g), "arrays")


if __name__ == "__main__":
    main()
