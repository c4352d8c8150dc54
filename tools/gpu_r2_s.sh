mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -x > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nccl.log; tail -5 gpurun_out/pytest_nccl.log | cut -c1-300
for mode in 1 0; do
NEMO_B200_DP_GRAPH_NCCL=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 1 --steps 40 --warmup 5 > gpurun_out/bench_n2_g$mode.json 2> gpurun_out/bench_n2_g$mode.err; cut -c1-220 gpurun_out/bench_n2_g$mode.json; tail -2 gpurun_out/bench_n2_g$mode.err | cut -c1-300
done
timeout 200 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_1of2.json 2>/dev/null; cut -c1-220 gpurun_out/bench_1of2.json
