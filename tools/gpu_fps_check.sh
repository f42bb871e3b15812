#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -x -k "fps or ops or smoke or model or c1 or pipeline" > gpurun_out/quick_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/quick_pytest.log
tail -4 gpurun_out/quick_pytest.log
timeout 100 python tools/fps_ab.py 2>&1 | tail -8
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; python -c "
import json; d=json.load(open('gpurun_out/quick_bench.json')); print(d['value'], d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'])"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"fps_" -s 2 -c 1 -f -o gpurun_out/fps_v3 python tools/fps_ab.py > gpurun_out/ncu_fps.log 2>&1
