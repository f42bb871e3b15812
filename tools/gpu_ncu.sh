#!/bin/bash
# ncu evidence for the current bench command: launch list (all kernels) + one full capture of a named kernel
# usage: bash tools/gpu_ncu.sh <kernel-regex> <tag>
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
CMD="python bench.py --steps 2 --warmup 1 --graph 0 --no-cpu-baseline --precision bf16x3"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$2.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$1 -s 4 -c 1 -o gpurun_out/prof_$2 $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8; tail -3 gpurun_out/ncu_full.log
