mkdir -p gpurun_out
for mode in 1 0; do
NEMO_B200_DP_SYMM=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 1 --steps 40 --warmup 5 > gpurun_out/bench_n2_symm$mode.json 2> gpurun_out/bench_n2_symm$mode.err; echo "rc=$?"; cut -c1-220 gpurun_out/bench_n2_symm$mode.json; grep -v "Warning\|_warn\|^$" gpurun_out/bench_n2_symm$mode.err | tail -4 | cut -c1-300
done
NEMO_B200_DP_SYMM=1 timeout 200 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -x -k "8" 2>&1 | tail -4 | cut -c1-300
