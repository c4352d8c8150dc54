set -x
mkdir -p gpurun_out
for f in 0 1 2 4 8 16; do NEMO_B200_GEMM_FLUSH=$f timeout 120 python tools/gemm_probe.py; done > gpurun_out/gemm_probe.txt 2>&1
cat gpurun_out/gemm_probe.txt
