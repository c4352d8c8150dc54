set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 120 python tools/bench_skin.py > gpurun_out/bench_skin.txt 2>&1; cat gpurun_out/bench_skin.txt
timeout 200 python tools/profile_step.py > gpurun_out/profile_step.txt 2>&1; cat gpurun_out/profile_step.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_b.log 2>&1
NCU_ONCE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_skin -o gpurun_out/k7_full -f python tools/bench_skin.py > gpurun_out/ncu_k7.log 2>&1
tail -3 gpurun_out/ncu_k7.log
