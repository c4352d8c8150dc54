# 2-GPU comparison of the shared-gradient reduction: none (dry), NCCL between two graphs (default), torch's NVLS multimem
# kernel in the graph (symm1), this library's peer-memory kernels in the graph (symm2).  bash tools/nccl_sweep.sh [N]
N=${1:-2}
run() { name=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 40 > gpurun_out/n${N}_$name.json 2> gpurun_out/n${N}_$name.err; python - <<PY
import json
for line in open('gpurun_out/n${N}_$name.json'):
    if line.startswith('{"metric'):
        d = json.loads(line)
        print('%-8s %d GPUs  device %.4f ms/step   e2e %.4f   e2e blocking %.4f' % ('$name', d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['blocking_ms_per_step']))
PY
}
run dry NEMO_B200_DP_DRY=1
run nccl A=1
run symm2 NEMO_B200_DP_SYMM=2


