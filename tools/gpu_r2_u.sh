mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_skin.py tests/test_gpu_gemm_tc.py -m gpu -q -x 2>&1 | tail -2
timeout 120 python tools/bench_skin.py 4096 2>&1 | head -4
timeout 120 python tools/bench_gemm.py 2>&1 | cut -c1-150
