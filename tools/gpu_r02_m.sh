#!/bin/bash
# round 2, GPU session m: chained activations after the group cursor (no 64-bit division per row) -- per-launch times chained / materialised, step time
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q -x --tb=short -k "chained or dz_formed or conv_bn or pool" > gpurun_out/r02m_pytest_chain.log 2>&1
echo "pytest chain rc=$?"; tail -5 gpurun_out/r02m_pytest_chain.log
timeout 300 python tools/train_kernel_times.py --out gpurun_out/r02m_train_kernel_times.json > gpurun_out/r02m_train_kernel_times.txt 2>&1
echo "times rc=$?"; cat gpurun_out/r02m_train_kernel_times.txt
timeout 300 python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_bench_train.json 2> gpurun_out/r02m_bench_train.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02m_bench_train.json'))
t=d.get('train', d)
print('ms_per_step', t['ms_per_step'], 'clouds/s', t['value'], 'launches/step', t.get('launches_per_step'))
PY
