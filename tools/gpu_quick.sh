#!/bin/bash
# usage: bash tools/gpu_quick.sh "<pytest -k expr>" [bench precision]
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout ${PYTEST_TIMEOUT:-300} python -m pytest tests -m gpu -q --tb=short -k "$1" > gpurun_out/quick_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/quick_pytest.log
timeout 120 python bench.py --steps 10 --warmup 3 --precision ${2:-bf16x3} --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
echo "rc=$?" >> gpurun_out/quick_bench.err
timeout 300 python tools/op_timing.py > gpurun_out/op_timing.log 2>&1
tail -25 gpurun_out/quick_pytest.log; python -c "
import json; d=json.load(open('gpurun_out/quick_bench.json')); print(d['value'], d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'])"; tail -2 gpurun_out/quick_bench.err; grep -E '^C[0-9]' gpurun_out/op_timing.log | cut -c1-400
