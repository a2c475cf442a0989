#!/usr/bin/env python
"""Micro-benchmark of the catalog scorers (fp32 FFMA vs bf16 tcgen05) on random P/Q (run under gpurun)."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_cbrs_amar_renaissance_b200 import ops

def main():
    n_users = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    n_items = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    P = torch.randn(n_users, 64, device=dev); Q = torch.randn(n_items, 64, device=dev)
    w2 = torch.randn(64, 64, device=dev) * 0.2; b2 = torch.randn(64, device=dev) * 0.1
    w3 = torch.randn(64, device=dev) * 0.2; b3 = torch.zeros(1, device=dev)
    for _ in range(2):
        ops.score_catalog_topk(P, Q, w2, b2, w3, b3, 10, precision=prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.score_catalog_topk(P, Q, w2, b2, w3, b3, 10, precision=prec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"precision": prec, "users": n_users, "items": n_items, "ms": ms, "pairs_per_s": n_users * n_items / ms * 1e3}))

main()
