set -x
mkdir -p gpurun_out
NEMO_B200_GEMM_TIMING=1 timeout 120 python tools/gemm_timeline.py > gpurun_out/gemm_timeline.txt 2>&1; cat gpurun_out/gemm_timeline.txt
timeout 600 python -m pytest tests/test_gpu_phases.py -q > gpurun_out/pytest_phases.log 2>&1; tail -5 gpurun_out/pytest_phases.log
