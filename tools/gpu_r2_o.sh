set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
NEMO_B200_DP_DRY=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 1 --steps 40 --warmup 5 > gpurun_out/bench_n2_dry.json 2> gpurun_out/bench_n2_dry.err; cut -c1-200 gpurun_out/bench_n2_dry.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 1 --steps 40 --warmup 5 > gpurun_out/bench_n2_c1.json 2> gpurun_out/bench_n2_c1.err; cut -c1-200 gpurun_out/bench_n2_c1.json
CUDA_VISIBLE_DEVICES=1 timeout 200 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_1of2.json 2>/dev/null; cut -c1-200 gpurun_out/bench_1of2.json
