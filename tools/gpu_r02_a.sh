#!/bin/bash
# round 2, GPU session a: the GPU test suite after the host-side changes, parity of the shipped precision at C3 / C4 / C5
# against the fp64 oracle (error distributions + attribution), and the tcgen05 instruction-shape microbenchmark
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02a_pytest_gpu.log
for c in C3 C4 C5; do
  timeout 600 python tools/parity_c3.py --config $c --out gpurun_out/r02a_parity_$c.json > gpurun_out/r02a_parity_$c.log 2>&1; echo "parity $c rc=$?"
done
timeout 120 python tools/umma_bench.py --cta-group 1 --out gpurun_out/r02a_umma_bench_cg1.json > gpurun_out/r02a_umma_cg1.log 2>&1; echo "umma cg1 rc=$?"
timeout 120 python tools/umma_bench.py --cta-group 2 --out gpurun_out/r02a_umma_bench_cg2.json > gpurun_out/r02a_umma_cg2.log 2>&1; echo "umma cg2 rc=$?"
tail -3 gpurun_out/r02a_umma_cg1.log gpurun_out/r02a_umma_cg2.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
