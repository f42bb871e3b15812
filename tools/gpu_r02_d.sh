#!/bin/bash
# round 2, GPU session d: det_rows with N = 128 pair MMAs -- parity tests first (own timeout: a protocol bug would hang), then timeline + bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -x -q > gpurun_out/r02d_pytest_tc.log 2>&1; rc=$?; echo "pytest tc rc=$rc"; tail -5 gpurun_out/r02d_pytest_tc.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_parity_gpu.py tests/test_nms_gpu.py -m gpu -x -q > gpurun_out/r02d_pytest_model.log 2>&1; echo "pytest model rc=$?"; tail -5 gpurun_out/r02d_pytest_model.log
timeout 120 python tools/tc_timeline.py > gpurun_out/r02d_det_timeline.txt 2>&1; echo "timeline rc=$?"; tail -4 gpurun_out/r02d_det_timeline.txt
timeout 300 python bench.py --steps 20 --warmup 3 --workload infer > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02d_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('sm_partition'))
for k in d['kernels']: print('   ', k['kernel'], round(k['ms'],4), round(k['achieved'],1), k['unit'], round(k['frac'],3))
PY
