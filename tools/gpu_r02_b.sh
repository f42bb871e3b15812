#!/bin/bash
# round 2, GPU session b: GPU suite with the new parity tests / pipelined step, then the bench (inference + nested training step)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02b_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r02b_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --pipelined 0 --workload infer --no-cpu-baseline > gpurun_out/r02b_bench_serial.json 2> gpurun_out/r02b_bench_serial.err; echo "bench serial rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02b_smoke.log
