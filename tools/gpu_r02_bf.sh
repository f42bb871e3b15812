#!/bin/bash
# round 2, GPU session bf: training converters with the bit-level hi/lo residual -- training tests, training bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -x -q > gpurun_out/r02bf_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r02bf_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python bench.py --steps 20 --warmup 3 --workload train --no-cpu-baseline > gpurun_out/r02bf_bench_train.json 2> gpurun_out/r02bf_err.txt; echo "bench rc=$?"; tail -3 gpurun_out/r02bf_err.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02bf_bench_train.json').read().strip().splitlines()[-1])
t=d.get('train', d)
print(t['ms_per_step'], t['value'], t['config'].get('serial_ms_per_step'), t['e2e']['value'])
for k in t['kernels']: print('   %-46s %2d %.4f ms %.0f %s %.3f' % (k['kernel'], k['launches'], k['ms'], k['achieved'], k['unit'], k['frac']))
PY
