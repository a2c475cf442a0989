#!/usr/bin/env python
"""Micro-benchmark of the chained tensor-core hybrid catalog scorer (cbrs_score_hybrid_topk_bf16) on random hoisted
tables (run under gpurun): users x items, CUDA events, one warm-up.  40,960 tensor FLOP per pair (4 products of 64 x 64
plus the 128 x 64 first classifier layer counted as two)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_cbrs_amar_renaissance_b200 import ops  # noqa: E402


def main():
    shapes = [(6040, 3706), (4736, 200000)] if len(sys.argv) < 3 else [(int(sys.argv[1]), int(sys.argv[2]))]
    dev = torch.device("cuda", 0)
    g = torch.Generator(dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    c = 64
    w = [r(c, c) * 0.2, r(c) * 0.1, r(c, c) * 0.2, r(c) * 0.1, r(2 * c, c) * 0.15, r(c) * 0.1, r(c, c) * 0.2, r(c) * 0.1, r(c) * 0.2,
         torch.zeros(1, device=dev)]
    for n_users, n_items in shapes:
        t = [r(n_users, c), r(n_items, c), r(n_users, c), r(n_items, c)]
        ops.score_hybrid_topk_bf16(*t, *w, 10)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.score_hybrid_topk_bf16(*t, *w, 10)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        pairs = n_users * n_items
        print(json.dumps({"kernel": "score_hybrid_tc_kernel", "users": n_users, "items": n_items, "ms": ms,
                          "pairs_per_s": pairs / ms * 1e3, "tensor_tflops": pairs * 40960 / ms / 1e9}), flush=True)


main()
