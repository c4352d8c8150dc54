set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_gemm_tc.py -q -x > gpurun_out/pytest_gemm2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gemm2.log; tail -5 gpurun_out/pytest_gemm2.log | cut -c1-250
timeout 100 python tools/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1; cut -c1-150 gpurun_out/bench_gemm.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
