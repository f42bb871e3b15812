#!/bin/bash
# round 2, GPU session h: full validation at HEAD (every -m gpu test, smoke, both bench arms), the BASELINE.md section-3 baselines,
# per-operator timing vs the reference kernels, and the C5 file-flow tool on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=20 --durations=8 > gpurun_out/r02h_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02h_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02h_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
echo "bench rc=$?" >> gpurun_out/r02h_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02h_bench_ref.json 2> gpurun_out/r02h_bench_ref.err
echo "ref rc=$?" >> gpurun_out/r02h_bench_ref.err
timeout 600 python tools/baselines.py > gpurun_out/r02h_baselines.log 2>&1
echo "baselines rc=$?" >> gpurun_out/r02h_baselines.log
timeout 600 python tools/op_timing.py > gpurun_out/r02h_op_timing.log 2>&1
echo "op_timing rc=$?" >> gpurun_out/r02h_op_timing.log
timeout 600 python tools/c5_kitti_multi.py --scans 8 > gpurun_out/r02h_c5_1gpu.log 2>&1
echo "c5 rc=$?" >> gpurun_out/r02h_c5_1gpu.log
tail -12 gpurun_out/r02h_pytest_gpu.log; tail -3 gpurun_out/r02h_smoke.log; cut -c1-600 gpurun_out/r02h_bench.json; tail -2 gpurun_out/r02h_bench.err
cut -c1-300 gpurun_out/r02h_bench_ref.json; tail -3 gpurun_out/r02h_baselines.log | cut -c1-1500; tail -2 gpurun_out/r02h_op_timing.log | cut -c1-300; tail -3 gpurun_out/r02h_c5_1gpu.log | cut -c1-800
