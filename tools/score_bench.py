"""Full-catalog BasicRS scoring + top-10 (cbrs_score_catalog_topk*) at the reference's classifier shape (64 -> 64 -> 1):
FFMA kernel, fp32-accurate 3xTF32 tensor-core kernel, bf16 tensor-core kernel.  CUDA-event timing on the launching stream.

    python tools/score_bench.py [users] [items]        # one JSON line per kernel"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def time_ms(fn, iters=3, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(n_users=4736, n_items=200000, c1=64, c2=64, k=10):
    from deep_cbrs_amar_renaissance_b200 import ops
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(1)
    P = torch.randn(n_users, c1, device=dev, generator=g)
    Q = torch.randn(n_items, c1, device=dev, generator=g)
    w2 = torch.randn(c1, c2, device=dev, generator=g) / c1 ** 0.5
    b2 = torch.randn(c2, device=dev, generator=g) * 0.1
    w3 = torch.randn(c2, device=dev, generator=g) / c2 ** 0.5
    b3 = torch.full((1,), 0.05, device=dev)
    out = []
    ref = None
    for precision in ("fp32-ffma", "fp32", "bf16"):
        ids, vals = ops.score_catalog_topk(P, Q, w2, b2, w3, b3, k, precision)
        ms = time_ms(lambda: ops.score_catalog_topk(P, Q, w2, b2, w3, b3, k, precision))
        line = {"kernel": {"fp32-ffma": "cbrs_score_catalog_topk (FFMA)", "fp32": "cbrs_score_catalog_topk_tf32x3 (3xTF32 tcgen05)",
                           "bf16": "cbrs_score_catalog_topk_bf16 (bf16 tcgen05)"}[precision],
                "users": n_users, "items": n_items, "c1": c1, "c2": c2, "k": k, "ms": ms,
                "pairs_per_s": n_users * n_items / (ms * 1e-3)}
        if ref is None:
            ref = (ids, vals)
        else:
            line["ids_equal_to_ffma"] = float((ids == ref[0]).float().mean().item())
            line["max_score_diff_vs_ffma"] = float((vals - ref[1]).abs().max().item())
        out.append(line)
    return out


if __name__ == "__main__":
    torch.cuda.set_device(0)
    nu = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
    ni = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    for line in run(nu, ni):
        print(json.dumps(line), flush=True)
