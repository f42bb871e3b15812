#!/bin/bash
# round 2, GPU session ba: desc_rows with E2 on a warpgroup of its own -- parity tests first (own timeout), timeline, inference bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -x -q > gpurun_out/r02ba_pytest_tc.log 2>&1; rc=$?; echo "pytest tc rc=$rc"; tail -5 gpurun_out/r02ba_pytest_tc.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 120 python tools/tc_timeline.py desc > gpurun_out/r02ba_desc_timeline.txt 2>&1; echo "timeline rc=$?"; tail -30 gpurun_out/r02ba_desc_timeline.txt
timeout 300 python bench.py --steps 20 --warmup 3 --workload infer --no-cpu-baseline > gpurun_out/r02ba_bench.json 2> gpurun_out/r02ba_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02ba_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ba_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('sm_partition'))
print(d.get('stage_ms'))
for k in d['kernels']: print('   ', k['kernel'], round(k['ms'],4), round(k['achieved'],1), k['unit'], round(k['frac'],3))
PY
