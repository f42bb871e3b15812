#!/bin/bash
# round 2, GPU session bd: descriptor tail epilogue -- tc tests, tail scaling, inference bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/r02bd_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r02bd_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 200 python tools/tail_scaling.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['clouds'], d['rounds_148'], {k.replace('_kernel','').replace('post_tc',''): v for k, v in d['us'].items()})
"
timeout 300 python bench.py --steps 20 --warmup 3 --workload infer --no-cpu-baseline > gpurun_out/r02bd_bench.json 2> gpurun_out/r02bd_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02bd_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02bd_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('sm_partition'))
print(d.get('stage_ms'))
for k in d['kernels']: print('   ', k['kernel'], round(k['ms'],4), round(k['achieved'],1), k['unit'], round(k['frac'],3))
PY
