# weak / strong scaling at $1 GPUs: every command bounded
set -x
G=$1
mkdir -p gpurun_out
for c in 1 2 3; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --config $c --steps 20 --warmup 5 > gpurun_out/bench_n${G}_c$c.json 2> gpurun_out/bench_n${G}_c$c.err; cut -c1-200 gpurun_out/bench_n${G}_c$c.json; tail -2 gpurun_out/bench_n${G}_c$c.err | cut -c1-200
done
