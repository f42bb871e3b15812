#!/bin/bash
# full round-end style validation: every -m gpu test, smoke, both bench arms, the training-step bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=20 --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?" >> gpurun_out/bench_ref.err
timeout 300 python tools/train_bench.py > gpurun_out/train_bench.log 2>&1
echo "train rc=$?" >> gpurun_out/train_bench.log
tail -12 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cut -c1-600 gpurun_out/bench.json; tail -2 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_ref.json; tail -2 gpurun_out/train_bench.log | cut -c1-400
