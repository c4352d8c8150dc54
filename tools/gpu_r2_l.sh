set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_gemm_tc.py -q -x > gpurun_out/pytest_gemm2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gemm2.log; tail -12 gpurun_out/pytest_gemm2.log | cut -c1-250
timeout 100 python tools/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1; cut -c1-150 gpurun_out/bench_gemm.txt
NEMO_B200_GEMM_TIMING=1 timeout 100 python tools/gemm_timeline.py 2>&1 | grep -E "N=1000 K=1000|N=20736" | head -4 | cut -c1-420
