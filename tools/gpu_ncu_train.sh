#!/bin/bash
# one ncu --set full capture of the large training-step kernels (second step of tools/train_bench.py); the report is
# summarised to CSV on the box because gpurun_out is capped at 64 MiB
mkdir -p gpurun_out
export F3D_TRAIN_GRAPH=0
CMD="python tools/train_bench.py"
timeout 120 $CMD > gpurun_out/train_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/train_plain.log; exit 1; }
tail -1 gpurun_out/train_plain.log | cut -c1-200
timeout 500 ncu --set full --clock-control none -k regex:"lin_tc|wgrad_tc_kernel|bn_bwd_apply|bn_bwd_reduce_kernel|bn_apply_kernel" -s 49 -c 26 -o /tmp/prof_train_$1 $CMD > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/ncu_train.log
ncu -i /tmp/prof_train_$1.ncu-rep --page raw --csv > gpurun_out/prof_train_$1_raw.csv 2>/dev/null
ls -la /tmp/prof_train_$1.ncu-rep gpurun_out/prof_train_$1_raw.csv
