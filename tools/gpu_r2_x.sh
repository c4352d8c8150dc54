mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q -x -k h16 2>&1 | tail -8 | cut -c1-200
timeout 120 python tools/bench_gemm.py 2>&1 | grep "pose-blend" | cut -c1-160
