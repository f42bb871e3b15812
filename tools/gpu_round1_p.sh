#!/bin/bash
# final evidence run of the round: op timing vs the reference kernels, KITTI flow, ncu launch list + full capture, bench
mkdir -p gpurun_out
timeout 300 python tools/op_timing.py > gpurun_out/op_timing_p.log 2>&1; echo "op_timing rc=$?"
timeout 300 python tools/w4_kitti.py > gpurun_out/w4_p.log 2>&1; echo "w4 rc=$?"
timeout 900 bash tools/gpu_ncu_all.sh p
ncu -i gpurun_out/prof_p.ncu-rep --page raw --csv > gpurun_out/prof_p_raw.csv 2>/dev/null
rm -f gpurun_out/prof_p.ncu-rep
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "bench rc=$?"
ls -la gpurun_out | tail -8
