#!/bin/bash
# round 2, GPU session r: fused local frames, no dz store for the detector's xyz layer, no materialised zero gradients
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x --tb=short -k "local_frames or sample_and_group" > gpurun_out/r02r_pytest_frames.log 2>&1
echo "pytest frames rc=$?"; tail -15 gpurun_out/r02r_pytest_frames.log
timeout 1200 python -m pytest tests/test_train_gpu.py tests/test_parity_gpu.py tests/test_model_gpu.py -q --tb=short > gpurun_out/r02r_pytest_train.log 2>&1
echo "pytest train rc=$?"; tail -8 gpurun_out/r02r_pytest_train.log
timeout 300 python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_bench_train.json 2> gpurun_out/r02r_bench_train.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02r_bench_train.json'))
t=d.get('train', d)
print('ms_per_step', t['ms_per_step'], 'clouds/s', t['value'], 'launches/step', t.get('launches_per_step'))
PY
timeout 300 python tools/train_graph_timeline.py --out gpurun_out/r02r_train_graph_timeline.txt | head -60
