set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_skin.py -m gpu -q -x > gpurun_out/pytest_skin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_skin.log
tail -15 gpurun_out/pytest_skin.log | cut -c1-300
NEMO_B200_SKIN_BWD=1 timeout 120 python tools/bench_skin.py > gpurun_out/bench_skin_g1.txt 2>&1; cat gpurun_out/bench_skin_g1.txt
timeout 120 python tools/bench_skin.py > gpurun_out/bench_skin_g2.txt 2>&1; cat gpurun_out/bench_skin_g2.txt
NEMO_B200_SKIN_TIMING=1 timeout 120 python tools/bench_skin.py > gpurun_out/bench_skin_timing.txt 2>&1; grep timing gpurun_out/bench_skin_timing.txt | tail -4
