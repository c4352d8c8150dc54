set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dp_nccl.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 200 python tools/profile_step.py > gpurun_out/profile_step.txt 2>&1; cat gpurun_out/profile_step.txt
for c in 0 2 3 4; do timeout 400 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; cut -c1-250 gpurun_out/bench_c$c.json; tail -3 gpurun_out/bench_c$c.err; done
