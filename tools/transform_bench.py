#!/usr/bin/env python
"""GCN transform Z = X W at config-5 scale (1.1e7 x 128 x 128 fp32): FFMA kernel vs the 3xTF32 tensor-core kernel.
Prints one JSON line per kernel: ms (CUDA events, best and mean of 5 after a warm-up), algorithmic bytes = rows * (k + n)
* 4 over time against the measured copy bandwidth, and the largest deviation between the two results."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_cbrs_amar_renaissance_b200 import ops  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return ms


def main():
    rows, k, n = int(os.environ.get("ROWS", 11_000_000)), 128, 128
    dev = torch.device("cuda", 0)
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    gen = torch.Generator(dev).manual_seed(0)
    x = torch.randn(rows, k, device=dev, generator=gen) * 0.05
    w = torch.randn(k, n, device=dev, generator=gen) / 11.3
    out_a, out_b = torch.empty(rows, n, device=dev), torch.empty(rows, n, device=dev)
    alg = rows * (k + n) * 4
    for name, fn, out in (("dense_fast_kernel (fp32 FFMA)", lambda: ops.dense(x, w, out=out_a), out_a),
                          ("dense_tf32x3_kernel (3xTF32 tcgen05, TMA)", lambda: ops.dense_tf32x3(x, w, out=out_b), out_b)):
        ms = timed(fn)
        best = min(ms)
        print(json.dumps({"kernel": name, "rows": rows, "k": k, "n": n, "ms": ms, "best_ms": best,
                          "algorithmic_gb": alg / 1e9, "gbps": alg / best / 1e6, "frac_of_hbm_peak": alg / best / 1e6 / peak,
                          "tflops_fp32_equivalent": 2.0 * rows * k * n / best / 1e9}), flush=True)
    scale = float(out_a.abs().max())
    print(json.dumps({"max_abs_dev_over_scale": float((out_a - out_b).abs().max()) / scale, "scale": scale}), flush=True)


if __name__ == "__main__":
    main()
