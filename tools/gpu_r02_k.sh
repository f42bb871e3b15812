#!/bin/bash
# round 2, GPU session k: the fused dz source of the pool-only training layers -- bit-identity test, training tests, step time on / off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q -x --tb=short -k "dz_formed or pool or conv_mid" > gpurun_out/r02k_pytest_dz.log 2>&1
echo "pytest dz rc=$?"; tail -15 gpurun_out/r02k_pytest_dz.log
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_parity_gpu.py -q --tb=short -k "not dz_formed" > gpurun_out/r02k_pytest_train.log 2>&1
echo "pytest train rc=$?"; tail -5 gpurun_out/r02k_pytest_train.log
timeout 300 python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_train.json 2> gpurun_out/r02k_bench_train.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02k_bench_train.json'))
t=d.get('train', d)
print('ms_per_step', t['ms_per_step'], 'clouds/s', t['value'])
for k in t['kernels']: print('  %-50s %2d %.4f ms  %.0f GB/s'%(k['kernel'],k['launches'],k['ms'],k['achieved']))
PY
