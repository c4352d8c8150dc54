mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_skin.py -m gpu -q -x > gpurun_out/pytest_skin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_skin.log
tail -6 gpurun_out/pytest_skin.log | cut -c1-300
(
echo "== default"; timeout 120 python tools/bench_skin.py 4096 bwd

for a in $ABL; do echo "== ablate $a"; NEMO_B200_SKIN_ABLATE=$a timeout 120 python tools/bench_skin.py 4096 "bwd  dV"; done
echo "== timing"; NEMO_B200_SKIN_TIMING=1 timeout 120 python tools/bench_skin.py 4096 "bwd  dV" 2>&1 | grep "timing\|trace"
) > gpurun_out/bwd2_ablate.txt 2>&1
grep -v "trace\|read-outs" gpurun_out/bwd2_ablate.txt
grep "trace" gpurun_out/bwd2_ablate.txt | grep "group  9\|group 10"
