#!/bin/bash
# bounded compute-sanitizer pass (memcheck): the smoke step of the fit engine and the ragged K7 cases
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
export NEMO_B200_GEMM_FLUSH=0
timeout 200 $CS --tool memcheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_memcheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_smoke.log
tail -4 gpurun_out/sanitize_memcheck_smoke.log
timeout 150 $CS --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_skin.py -q -x -k "ragged and (1 or 7 or 37)" > gpurun_out/sanitize_memcheck_skin.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_skin.log
tail -4 gpurun_out/sanitize_memcheck_skin.log
