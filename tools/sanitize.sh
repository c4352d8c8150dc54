#!/bin/bash
# compute-sanitizer over the hot path on the GPU box (memcheck on the kernel-level tests and one fit step, racecheck and
# synccheck on the shared-memory / mbarrier heavy kernels through the smoke step).  Slow (10-50x): run on demand,
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Output: gpurun_out/sanitize_*.log; the summary lines ("ERROR SUMMARY: 0 errors") are copied to profiles/.
set -x
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
export NEMO_B200_GEMM_FLUSH=0
timeout 600 $CS --tool memcheck --launch-timeout 0 --error-exitcode 1 python -m pytest tests/test_gpu_skin.py tests/test_gpu_gemm_tc.py -q -x -k "not 4096 and not full" > gpurun_out/sanitize_memcheck_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_kernels.log
tail -5 gpurun_out/sanitize_memcheck_kernels.log
timeout 400 $CS --tool memcheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_memcheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_smoke.log
tail -4 gpurun_out/sanitize_memcheck_smoke.log
timeout 400 $CS --tool racecheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_racecheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_racecheck_smoke.log
tail -4 gpurun_out/sanitize_racecheck_smoke.log
timeout 300 $CS --tool synccheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_synccheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_synccheck_smoke.log
tail -4 gpurun_out/sanitize_synccheck_smoke.log
