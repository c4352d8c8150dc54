# K7 backward ablation sweep (debug switches of csrc/skin_tc.cu); output: gpurun_out/ablate.txt
mkdir -p gpurun_out; : > gpurun_out/ablate.txt
for abl in ${ABLS:-0 1 2 4 8 16 26 31}; do
  echo "== ABLATE=$abl" >> gpurun_out/ablate.txt
  NEMO_B200_SKIN_ABLATE=$abl timeout 120 python tools/bench_skin.py 2>&1 | grep "bwd  dV" >> gpurun_out/ablate.txt
done
for c in ${CHUNKS:-}; do echo "== CHUNKS=$c" >> gpurun_out/ablate.txt; NEMO_B200_SKIN_CHUNKS=$c timeout 120 python tools/bench_skin.py 2>&1 | grep -E "fwd  V|l1 |bwd  dV" >> gpurun_out/ablate.txt; done
echo "== TIMING" >> gpurun_out/ablate.txt
NEMO_B200_SKIN_TIMING=1 timeout 120 python tools/bench_skin.py 2>&1 | grep -E "timing" | tail -5 >> gpurun_out/ablate.txt
cat gpurun_out/ablate.txt
