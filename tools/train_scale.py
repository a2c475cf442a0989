#!/usr/bin/env python
"""One optimiser step of BasicGCN (3 x 128, concatenation, BasicRS [48,48]/[64,64]) on a scaled synthetic graph: shows the
explicit backward kernels beyond MovieLens size.  Prints one JSON line (ms per step, loss trajectory on a fixed batch)."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from deep_cbrs_amar_renaissance_b200 import ops  # noqa: E402
from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph  # noqa: E402
from deep_cbrs_amar_renaissance_b200.keras_like import set_seed  # noqa: E402
from deep_cbrs_amar_renaissance_b200.models import basic  # noqa: E402


def main():
    scale = sys.argv[1] if len(sys.argv) > 1 else "c5-tenth"
    n_users, n_items, n_edges = {"c5-tenth": (1_000_000, 100_000, 100_000_000), "c5-hundredth": (100_000, 10_000, 10_000_000),
                                 "c5-fifth": (2_000_000, 200_000, 200_000_000)}[scale]
    batch = 65536
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    row, col = ops.synth_bipartite(n_users, n_items, n_edges, 42, dev)
    graph = DeviceGraph(row, col, None, n_users + n_items)
    del row, col
    set_seed(42)
    model = basic.BasicGCN(graph, n_hiddens=[128] * 3, embedding_dim=128, dense_units=[48, 48], clf_units=[64, 64],
                           final_node="concatenation", l2_regularizer=1e-6)
    nnz = graph.norm.nnz
    graph.release_coo()
    model.compile(loss="binary_crossentropy", optimizer={"learning_rate": 1e-3})
    rng = np.random.RandomState(0)
    u = rng.randint(0, n_users, size=batch)
    i = rng.randint(0, n_items, size=batch) + n_users
    y = ((u + i) % 2).astype(np.int64)  # a learnable parity rule
    ud, idv = torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev)
    losses = []
    for _ in range(3):
        loss, _ = model.train_on_batch((ud, idv), y)
        losses.append(float(loss.item()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    ops.LAUNCHES = 0
    e0.record()
    for _ in range(steps):
        loss, correct = model.train_on_batch((ud, idv), y)
    e1.record()
    torch.cuda.synchronize()
    losses.append(float(loss.item()))
    ms = e0.elapsed_time(e1) / steps
    # forward + backward each run 3 sparse passes over nnz edges
    print(json.dumps({"model": "BasicGCN 3x128", "scale": scale, "nodes": n_users + n_items, "nnz_a_hat": nnz, "batch": batch,
                      "train_step_ms": ms, "launches_per_step": ops.LAUNCHES / steps,
                      "edges_per_s_fwd_plus_bwd": 6 * nnz / (ms * 1e-3), "loss_first_3": losses[:3], "loss_after_13": losses[-1],
                      "accuracy_last": float(correct.item()) / batch}), flush=True)
    assert np.isfinite(losses).all() and losses[-1] < losses[0]


if __name__ == "__main__":
    main()
