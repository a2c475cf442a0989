"""Tower GEMMs of the hybrid scorer (econfigs/hybrid-gnn.yaml grid 2: Dense 768 -> 256 -> 64 over BERT rows) on the fp32
FFMA kernel (cbrs_dense) and on the tensor cores (cbrs_dense_tc), timed with CUDA events on the launching stream.

    python tools/dense_tc_bench.py [rows]        # one JSON line per shape on stdout

Inputs are larger than L2 (rows x 768 fp32 = 3.2 GB at the default 2**20 rows).  Rooflines: tensor = 2*m*k*n flop vs the
measured bf16 peak; hbm = algorithmic bytes (m*k*4 read + m*n*4 written + the bf16 image once) vs the measured copy
bandwidth - the layer is HBM-bound on the fp32 rows it reads (arithmetic intensity 2*n/4 = 128 flop/B at n = 256)."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fp:
            p = json.load(fp)
        return float(p.get("hbm_gbs") or 6544.3), float(p.get("bf16_tflops_sustained") or 1398.0)
    except Exception:
        return 6544.3, 1398.0


def time_ms(fn, iters=10, warmup=3):
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


def run(rows=1 << 20, shapes=((768, 256), (256, 64)), gather=False):
    from deep_cbrs_amar_renaissance_b200 import ops
    hbm, tflops = peaks()
    dev = torch.device("cuda", torch.cuda.current_device())
    out = []
    for k, n in shapes:
        x = torch.randn(rows, k, device=dev) * 0.5
        w = torch.randn(k, n, device=dev) / k ** 0.5
        b = torch.randn(n, device=dev) * 0.1
        y = torch.empty(rows, n, device=dev)
        idx = torch.randperm(rows, device=dev) if gather else None
        image = ops.dense_tc_image(w)
        ms_tc = time_ms(lambda: ops.dense_tc(x, w, b, "relu", idx1=idx, out=y, image=image))
        ms_fp = time_ms(lambda: ops.dense(x, w, b, "relu", idx1=idx, out=y), iters=3, warmup=1)
        flop = 2.0 * rows * k * n
        byt = rows * (k + n) * 4.0 + image.numel()
        tma = None
        if ops.dense_tc_bf16_eligible(k, 0, n):
            # the same layer over the table stored as bf16 once, rows fed by TMA (cbrs_dense_tc_bf16): fp32 and bf16 output
            xb = ops.to_bf16(x)
            yb = torch.empty(rows, n, device=dev, dtype=torch.bfloat16)
            ms_f = time_ms(lambda: ops.dense_tc_bf16(xb, w, b, "relu", idx1=idx, out=y, image=image))
            ms_b = time_ms(lambda: ops.dense_tc_bf16(xb, w, b, "relu", idx1=idx, out=yb, image=image))
            byt_f, byt_b = rows * (k * 2.0 + n * 4.0) + image.numel(), rows * (k + n) * 2.0 + image.numel()
            tma = {"out_f32": {"ms": ms_f, "tflops": flop / ms_f / 1e9, "frac_of_bf16_peak": flop / ms_f / 1e9 / tflops,
                               "hbm_gbps": byt_f / ms_f / 1e6, "frac_of_hbm_peak": byt_f / ms_f / 1e6 / hbm},
                   "out_bf16": {"ms": ms_b, "tflops": flop / ms_b / 1e9, "frac_of_bf16_peak": flop / ms_b / 1e9 / tflops,
                                "hbm_gbps": byt_b / ms_b / 1e6, "frac_of_hbm_peak": byt_b / ms_b / 1e6 / hbm},
                   "bytes": "rows as bf16 (k * 2 B) read once + output once + the image once"}
            del xb, yb
        out.append({"op": "dense %d->%d relu" % (k, n), "rows": rows, "gathered": bool(gather),
                    "tc_bf16": {"ms": ms_tc, "tflops": flop / ms_tc / 1e9, "frac_of_bf16_peak": flop / ms_tc / 1e9 / tflops,
                                "hbm_gbps": byt / ms_tc / 1e6, "frac_of_hbm_peak": byt / ms_tc / 1e6 / hbm},
                    "fp32_ffma": {"ms": ms_fp, "tflops": flop / ms_fp / 1e9}, "tc_bf16_table_tma": tma,
                    "speedup": ms_fp / ms_tc, "peaks": {"hbm_gbps": hbm, "bf16_tflops": tflops}})
        del x, y
    return out


if __name__ == "__main__":
    torch.cuda.set_device(0)
    n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    for g in (False, True):
        for line in run(n_rows, gather=g):
            print(json.dumps(line), flush=True)
