#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -x -k "tc or model or smoke or pipeline or c1" > gpurun_out/quick_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/quick_pytest.log
tail -3 gpurun_out/quick_pytest.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; python -c "
import json; d=json.load(open('gpurun_out/quick_bench.json')); print(d['value'], d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['fp32_path']['max_rel_attention_diff'], d['fp32_path']['max_abs_descriptor_diff'])"
timeout 100 python tools/tc_timeline.py desc 2>&1 | tail -16
timeout 100 python tools/tc_timeline.py det 2>&1 | tail -14
