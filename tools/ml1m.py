#!/usr/bin/env python
"""Latency of the per-batch model call at MovieLens-1M shape (BASELINE configs 1-4), eager vs
CUDA-graph replay, next to the CPU oracle.  These shapes are L2-resident and launch-bound: the
numbers are reported as time, never as an HBM fraction (run under gpurun; writes JSON lines)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_cbrs_amar_renaissance_b200.graphed import GraphedForward  # noqa: E402
from deep_cbrs_amar_renaissance_b200.keras_like import set_seed  # noqa: E402
from deep_cbrs_amar_renaissance_b200.models import basic, hybrid  # noqa: E402
from tests.helpers import random_bipartite  # noqa: E402


def timeit(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def cpu_train_step_ms(name, model, adj, inputs, y):
    """the oracle's autograd twin (float32, all host threads): forward + backward of one batch"""
    from oracle import graph as og
    from oracle import train as ot
    from tests.helpers import export_weights
    kind = {"BasicGCN": "gcn", "BasicGraphSage": "sage", "BasicLightGCN": "lightgcn", "BasicGAT": "gat"}.get(name)
    if kind is None:
        return None
    w = export_weights(model)
    if kind in ("sage", "gat"):
        ptr, idx, _ = og.reorder_raw(adj)
        graph = (ptr, idx)
    else:
        graph = og.gcn_filter(adj)
    fn = "mean" if kind == "lightgcn" else "concatenation"
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        loss, _, _, leaves = ot.forward_loss(kind, w, graph, inputs, y, final_node=fn, l2=1e-4, dtype=torch.float32)
        loss.backward()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


def main():
    torch.cuda.set_device(0)
    n_users, n_items, batch = 6040, 3706, 2048
    adj = random_bipartite(n_users, n_items, 572000, seed=42)
    adj_uip = random_bipartite(n_users, n_items, 572000, seed=42, n_props=17554, n_links=70341, dup_links=400)
    rng = np.random.RandomState(0)
    u = rng.randint(0, n_users, size=batch)
    i = rng.randint(0, n_items, size=batch) + n_users
    bert = (rng.standard_normal((n_users + n_items, 768)) * 0.5).astype(np.float32)
    cases = [("BasicGCN", basic, adj, {}), ("BasicGraphSage", basic, adj, {}), ("BasicGAT", basic, adj, {}),
             ("BasicLightGCN", basic, adj, {}), ("BasicDGCF", basic, adj, {}), ("BasicGCN-uip", basic, adj_uip, {}), ("BasicGAT-uip", basic, adj_uip, {}),
             ("HybridBertGCN", hybrid, adj, dict(dense_units=[[48, 48], [256, 64], [64, 64]], feature_based=True))]
    for name, mod, a, extra in cases:
        set_seed(42)
        kw = dict(n_hiddens=[16, 16], n_layers=2, embedding_dim=16, dense_units=[48, 48], clf_units=[64, 64])
        kw.update(extra)
        model = getattr(mod, name.split("-")[0])(a, **kw)
        is_h = mod is hybrid
        if is_h:
            model.set_content_table(bert)
        inputs = (u, i)
        model(inputs)
        gph = model.gnn.gnn_layers.adj_matrix
        nnz = (gph.raw if ("Sage" in name or "GAT" in name) else gph.dgcf if "DGCF" in name else gph.norm).nnz
        eager = timeit(lambda: model(inputs))
        g = GraphedForward(model, batch)
        graphed = timeit(lambda: g(inputs))
        assert torch.equal(g(inputs), model(inputs))
        model.cache_propagation = True
        model.propagate()
        cat = timeit(lambda: model.recommend_top_k(n_users, n_items, 10), n=5, warm=1) if not is_h else None
        model.cache_propagation = False
        model.invalidate()
        # one optimiser step (forward with saved activations, explicit backward, BCE + l2, Adam), batch 1024 as in
        # config.yaml:46; next to the same step on the host cores (torch-CPU autograd twin, fp32, all threads)
        train_ms = cpu_train_ms = train_graph_ms = None
        if True:
            yb = rng.randint(0, 2, size=1024)
            ub, ib = u[:1024], i[:1024]
            model.compile(loss="binary_crossentropy", optimizer={"learning_rate": 1e-3})
            train_graph_ms = None
            try:
                train_ms = timeit(lambda: model.train_on_batch((ub, ib), yb), n=20, warm=3)
                if True:
                    gstep = model.make_graphed_train_step(1024)
                    train_graph_ms = timeit(lambda: gstep((ub, ib), yb), n=50, warm=3)
            except NotImplementedError:
                train_ms = None
            if train_ms is not None and not is_h and "uip" not in name:
                cpu_train_ms = cpu_train_step_ms(name, model, a, (ub, ib), yb)
        print(json.dumps({"model": name, "nnz": nnz, "layers": 2, "batch": batch, "eager_ms": eager, "graph_ms": graphed,
                          "train_step_ms_batch1024": train_ms, "train_step_graph_ms_batch1024": train_graph_ms, "cpu_train_step_ms_batch1024": cpu_train_ms,
                          "cpu_threads": torch.get_num_threads(),
                          "edges_per_s_graph": 2 * nnz / (graphed * 1e-3),
                          "catalog_top10_ms": cat, "catalog_pairs_per_s": (n_users * n_items / (cat * 1e-3)) if cat else None}),
              flush=True)


if __name__ == "__main__":
    main()
