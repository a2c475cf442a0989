#!/usr/bin/env bash
# compute-sanitizer over the kernel-level GPU tests: memcheck on every kernel family, racecheck on
# the kernels that hand shared memory between warps / proxies (mbarrier rings, TMEM, bulk copies).  Run under gpurun:
#   gpurun --timeout 2400 -- 'bash tools/sanitize.sh'
# Logs land in gpurun_out/r02_sanitizer_*.log; the summaries are copied to profiles/.
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
MEM_TESTS="tests/test_gpu_kernels.py tests/test_gpu_blocked.py tests/test_gpu_tf32.py tests/test_gpu_hybrid_catalog.py tests/test_zz_gpu_dense_tc.py"
timeout 1500 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 --launch-timeout 0 \
    python -m pytest $MEM_TESTS -x -q -k "not radix_sort_is_stable and not each_kernel_variant" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -6 gpurun_out/r02_sanitizer_memcheck.log
RACE_TESTS="tests/test_gpu_tf32.py tests/test_gpu_hybrid_catalog.py tests/test_zz_gpu_dense_tc.py tests/test_gpu_blocked.py"
timeout 900 $SAN --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest $RACE_TESTS tests/test_gpu_kernels.py -x -q -k "tf32x3_matches_float64_product or chained_hybrid or single_source or fused_gcn or tensor_core_catalog or fused_catalog or spmm_blocked_matches" \
    > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log
tail -6 gpurun_out/r02_sanitizer_racecheck.log
