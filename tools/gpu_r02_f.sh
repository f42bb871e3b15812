#!/bin/bash
# round 2, GPU session f (2 GPUs): the bench (inference + nested training step with the NCCL all-reduce inside the graph) as the driver launches it
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02f_bench_${N}gpu.json 2> gpurun_out/r02f_bench_${N}gpu.err; echo "bench $N rc=$?"; tail -5 gpurun_out/r02f_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02f_bench_${N}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
t=d['train']; print('train', t['n_gpus'], t['ms_per_step'], t['value'], 'e2e', t['e2e']['ms_per_step'], t['allreduce'])
PY
