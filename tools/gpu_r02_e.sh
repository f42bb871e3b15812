#!/bin/bash
# round 2, GPU session e: desc_rows with pair MMAs + two producer warpgroups -- parity tests first (own timeout), timelines, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -x -q > gpurun_out/r02e_pytest_tc.log 2>&1; rc=$?; echo "pytest tc rc=$rc"; tail -5 gpurun_out/r02e_pytest_tc.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_tc_gpu.py > gpurun_out/r02e_pytest_rest.log 2>&1; echo "pytest rest rc=$?"; tail -5 gpurun_out/r02e_pytest_rest.log
timeout 120 python tools/tc_timeline.py desc > gpurun_out/r02e_desc_timeline.txt 2>&1; echo "timeline rc=$?"; tail -3 gpurun_out/r02e_desc_timeline.txt
timeout 300 python bench.py --steps 20 --warmup 3 --workload infer > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('sm_partition'))
for k in d['kernels']: print('   ', k['kernel'], round(k['ms'],4), round(k['achieved'],1), k['unit'], round(k['frac'],3))
PY
