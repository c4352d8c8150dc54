# final multi-GPU evidence: weak scaling (config 1) and configs[3] at $1 GPUs, every command bounded
G=$1
mkdir -p gpurun_out
for c in 1 3; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --config $c --steps 20 --warmup 5 > gpurun_out/bench_final_n${G}_c$c.json 2> gpurun_out/bench_final_n${G}_c$c.err; grep -o '"value": [0-9.]*, "unit": "frames/s", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' gpurun_out/bench_final_n${G}_c$c.json | head -1; tail -1 gpurun_out/bench_final_n${G}_c$c.err | cut -c1-200
done
