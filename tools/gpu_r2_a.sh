# round 2, call A: new parity tests + GEMM profile + bench on the current kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dp_nccl.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 200 python tools/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1; cat gpurun_out/bench_gemm.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
NCU_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -c 14 -o gpurun_out/gemm_full -f python tools/bench_gemm.py > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log
ls -la gpurun_out
