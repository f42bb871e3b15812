#!/bin/bash
# round 2, GPU session bi (one 8 x B200 box): bench.py (inference + training step with the NCCL all-reduce inside the graph) at 8 ranks, end of round 2
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; }
timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 20 --warmup 3" > gpurun_out/r02bi_bench_8gpu.json 2> gpurun_out/r02bi_bench_8gpu.err
echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02bi_bench_8gpu.json'))
print('infer', d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k not in ('how','copies_alone_note')})
t=d['train']; print('train', t['value'], t['ms_per_step'], t['allreduce'], t['e2e']['ms_per_step'])
PY
tail -3 gpurun_out/r02bi_bench_8gpu.err
