#!/usr/bin/env bash
# One GPU call that validates what was written after round 1's last GPU session:
#   gpurun --timeout 300 -- 'bash tools/experimental_check.sh'
# 1. the opt-in tests (relational backward, experimental Dense variant), 2. the tower bench with the shipped kernel and
# with CBRS_DENSE_TC_VARIANT=4.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
CBRS_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_zz_gpu_dense_tc.py tests/test_zz_gpu_kg_training.py -q -x \
    -k "experimental or rgcn" > gpurun_out/experimental_tests.log 2>&1
tail -5 gpurun_out/experimental_tests.log
timeout 60 python tools/dense_tc_bench.py > gpurun_out/dense_tc_bench_shipped.jsonl 2> gpurun_out/dense_tc_bench_shipped.err
CBRS_DENSE_TC_VARIANT=4 timeout 60 python tools/dense_tc_bench.py > gpurun_out/dense_tc_bench_variant4.jsonl 2> gpurun_out/dense_tc_bench_variant4.err
cut -c1-220 gpurun_out/dense_tc_bench_shipped.jsonl gpurun_out/dense_tc_bench_variant4.jsonl
