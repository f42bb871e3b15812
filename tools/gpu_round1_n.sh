#!/bin/bash
# evidence run after the FPS / tensor-kernel changes: bench line, op timing vs the reference kernels, KITTI flow, ncu
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo "bench rc=$?"
timeout 300 python tools/op_timing.py > gpurun_out/op_timing_n.log 2>&1; echo "op_timing rc=$?"
timeout 300 python tools/w4_kitti.py > gpurun_out/w4_n.log 2>&1; echo "w4 rc=$?"
timeout 900 bash tools/gpu_ncu_all.sh n
ncu -i gpurun_out/prof_n.ncu-rep --page raw --csv > gpurun_out/prof_n_raw.csv 2>/dev/null
rm -f gpurun_out/prof_n.ncu-rep
ls -la gpurun_out | tail -12
