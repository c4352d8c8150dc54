# K7 forward, second generation: correctness (skin tests + full-size properties) and timing against the first
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_skin.py tests/test_gpu_fullsize.py -q -x > gpurun_out/pytest_skin.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_skin.log; tail -15 gpurun_out/pytest_skin.log
timeout 120 python tools/bench_skin.py > gpurun_out/bench_skin2.txt 2>&1; cat gpurun_out/bench_skin2.txt
NEMO_B200_SKIN_FWD=1 timeout 120 python tools/bench_skin.py > gpurun_out/bench_skin1.txt 2>&1; head -3 gpurun_out/bench_skin1.txt
