mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_skin.py -m gpu -q -x 2>&1 | tail -2
timeout 120 python tools/bench_skin.py 4096 2>&1 | head -4
NCU_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_skin -o gpurun_out/k7_r2 -f python tools/bench_skin.py 4096 > gpurun_out/ncu_k7_r2.log 2>&1; tail -3 gpurun_out/ncu_k7_r2.log
ls -la gpurun_out/k7_r2.ncu-rep
