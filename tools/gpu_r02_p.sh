#!/bin/bash
# round 2, GPU session p: lin_tc epilogue with pointer bumps (no 64-bit product / bounds test per element) -- parity, phases, trace, step time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -x --tb=short > gpurun_out/r02p_pytest_train.log 2>&1
echo "pytest train rc=$?"; tail -5 gpurun_out/r02p_pytest_train.log
timeout 300 python tools/lin_tc_phases.py --out gpurun_out/r02p_lin_tc_phases.json 2>&1 | grep -A1 "MB, HBM" 
timeout 300 python tools/lin_tc_trace.py 2>&1 | grep -A2 "mask 0"
timeout 300 python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02p_bench_train.json 2> gpurun_out/r02p_bench_train.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02p_bench_train.json'))
t=d.get('train', d)
print('ms_per_step', t['ms_per_step'], 'clouds/s', t['value'], 'launches/step', t.get('launches_per_step'))
for k in t['kernels']: print('  %-50s %2d %.4f ms  %.0f GB/s'%(k['kernel'],k['launches'],k['ms'],k['achieved']))
PY
