set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_gemm_tc.py -q -x > gpurun_out/pytest_gemm2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gemm2.log; tail -30 gpurun_out/pytest_gemm2.log | cut -c1-250
