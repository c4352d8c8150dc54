set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -x > gpurun_out/pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gemm.log; tail -15 gpurun_out/pytest_gemm.log
timeout 200 python tools/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1; cat gpurun_out/bench_gemm.txt
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dp_nccl.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
