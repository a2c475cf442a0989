#!/usr/bin/env python
"""Sweep of the column-blocked schedule (graph.blocking_policy) on the config-5 graph, one GPU.

    python tools/tune_blocked.py [--scale c5] [--cols 32768,65536,...] [--min-len 512,1024,2048]

Prints one JSON line per setting: sparse-kernel time for a 128-wide fp32 operand (CUDA events, 3 launches after one
warm-up), chunk / slot counts, and the largest deviation from the row-major schedule's result (the reduction tree of
blocked rows differs, so the results agree to rounding, not bit for bit).  Also times the user-row half and the
item-row half of the row-major launch separately: they gather from tables of very different size (512 MB vs 5.1 GB).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from bench import SCALES  # noqa: E402
from deep_cbrs_amar_renaissance_b200 import ops  # noqa: E402
from deep_cbrs_amar_renaissance_b200.graph import CsrSlice, DeviceGraph  # noqa: E402


def time_spmm(csr, x, out, reps=3, agg=0):
    ops.spmm(csr, x, out, agg=agg)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for r in range(reps):
        ops.spmm(csr, x, out, agg=agg)
        ev[r + 1].record()
    torch.cuda.synchronize()
    return [ev[r].elapsed_time(ev[r + 1]) for r in range(reps)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", default="c5")
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--cols", default="32768,49152,65536,98304,131072,196608")
    ap.add_argument("--min-len", default="512,1024,2048")
    ap.add_argument("--halves", action="store_true", help="also time user rows and item rows separately (row-major)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n_users, n_items, n_edges = SCALES[args.scale]
    n = n_users + n_items
    os.environ["CBRS_BLOCK_COLS"] = "0"  # the baseline view is row-major
    row, col = ops.synth_bipartite(n_users, n_items, n_edges, 42, dev)
    g = DeviceGraph(row, col, None, n)
    del row, col
    base = g.norm
    g.release_coo()
    torch.cuda.empty_cache()
    x = torch.randn(n, args.dim, device=dev, generator=torch.Generator(dev).manual_seed(0)) * 0.05
    ref = torch.empty(n, args.dim, device=dev)
    ms = time_spmm(base, x, ref)
    scale = float(ref.abs().max())
    print(json.dumps({"schedule": "row-major", "ms": ms, "n_chunks": base.chunks["n_chunks"], "n_heavy": base.chunks["n_heavy"],
                      "n_slots": base.chunks["n_slots"], "nnz": base.nnz}), flush=True)
    if args.halves:
        for name, (r0, r1) in (("user rows", (0, n_users)), ("item rows", (n_users, n))):
            s = base.row_slice(r0, r1)
            out = torch.empty(r1 - r0, args.dim, device=dev)
            print(json.dumps({"schedule": "row-major, " + name, "ms": time_spmm(s, x, out), "nnz": s.nnz}), flush=True)
            del s, out
        torch.cuda.empty_cache()
    out = torch.empty(n, args.dim, device=dev)
    for cols in (int(c) for c in args.cols.split(",")):
        for min_len in (int(m) for m in args.min_len.split(",")):
            csr = CsrSlice(base.rowptr, base.colidx, base.vals, base.n_cols, base.chunk_edges, blocking=(cols, min_len))
            ms = time_spmm(csr, x, out)
            err = float((out - ref).abs().max()) / scale
            print(json.dumps({"schedule": "blocked", "block_cols": cols, "block_min_len": min_len, "ms": ms,
                              "window_mb": cols * args.dim * 4 / 2 ** 20, "n_chunks": csr.chunks["n_chunks"],
                              "n_heavy": csr.chunks["n_heavy"], "n_slots": csr.chunks["n_slots"],
                              "max_rel_dev_vs_row_major": err}), flush=True)
            del csr
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
