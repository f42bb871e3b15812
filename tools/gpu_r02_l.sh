#!/bin/bash
# round 2, GPU session l: chained (unmaterialised) activations of the training layers -- bit-identity test, training tests, step time
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q -x --tb=short -k "chained or dz_formed" > gpurun_out/r02l_pytest_chain.log 2>&1
echo "pytest chain rc=$?"; tail -25 gpurun_out/r02l_pytest_chain.log
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_parity_gpu.py -q --tb=short -k "not chained and not dz_formed" > gpurun_out/r02l_pytest_train.log 2>&1
echo "pytest train rc=$?"; tail -5 gpurun_out/r02l_pytest_train.log
timeout 300 python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench_train.json 2> gpurun_out/r02l_bench_train.err
echo "bench rc=$?"; tail -3 gpurun_out/r02l_bench_train.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench_train.json'))
t=d.get('train', d)
print('ms_per_step', t['ms_per_step'], 'clouds/s', t['value'], 'launches/step', t.get('launches_per_step'))
for k in t['kernels']: print('  %-50s %2d %.4f ms  %.0f GB/s'%(k['kernel'],k['launches'],k['ms'],k['achieved']))
PY
