#!/bin/bash
# round 2, GPU session bp: full GPU suite + smoke + default bench at HEAD (after the NMS / ball-query pass)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02bp_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?"; tail -8 gpurun_out/r02bp_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02bp_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/r02bp_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02bp_bench.json 2> gpurun_out/r02bp_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02bp_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02bp_bench.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
print('stage', d.get('stage_ms'))
for k in d['kernels']: print('  %-34s %2d %.4f ms  %.1f %s'%(k['kernel'],k['launches'],k['ms'],k['achieved'],k['unit']))
t=d.get('train')
if t: print('train', t['ms_per_step'], t['value'], t['e2e'])
print('cpu', d.get('cpu_baseline'))
PY
