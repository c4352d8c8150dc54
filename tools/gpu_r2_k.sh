# data-parallel graphs (NCCL outside): 2-rank NCCL parity + 2-GPU bench; every command bounded
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -x > gpurun_out/pytest_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_nccl.log; tail -25 gpurun_out/pytest_nccl.log
for c in 1 2 4; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $c --steps 20 --warmup 5 > gpurun_out/bench_n2_c$c.json 2> gpurun_out/bench_n2_c$c.err; cut -c1-200 gpurun_out/bench_n2_c$c.json; tail -4 gpurun_out/bench_n2_c$c.err | cut -c1-300
done
