#!/bin/bash
# round 2, GPU session bu: index windows sized to powers of two (every cloud of >= 32768 points on the grouped ball query) -- full GPU suite, smoke,
# W4 flow, default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02bu_pytest_gpu.log 2>&1
rc=$?; echo "pytest gpu rc=$rc"; tail -4 gpurun_out/r02bu_pytest_gpu.log
if [ $rc -ne 0 ]; then grep -n "Error\|FAILED" gpurun_out/r02bu_pytest_gpu.log | head -5; exit 0; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02bu_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r02bu_smoke.log
timeout 300 python tools/w4_kitti.py > gpurun_out/r02bu_w4.jsonl 2> gpurun_out/r02bu_w4.err; echo "w4 rc=$?"; grep bf16x3 gpurun_out/r02bu_w4.jsonl | cut -c1-330
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02bu_bench.json 2> gpurun_out/r02bu_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r02bu_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02bu_bench.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
t=d.get('train')
if t: print('train', t['ms_per_step'], t['value'])
PY
