#!/bin/bash
# round 2, GPU session g: ncu evidence -- launch list + one --set full capture of every hot inference kernel (serial eager step), and one
# --set full capture of the large training-step kernels (second eager step).  Raw CSVs come back; summaries are made by tools/ncu_summary.py.
mkdir -p gpurun_out
CMD="python bench.py --workload infer --steps 2 --warmup 1 --graph 0 --pipelined 0 --no-cpu-baseline"
$CMD > gpurun_out/r02g_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches_infer.csv $CMD > gpurun_out/r02g_ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r02g_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fps_group|bq_grid|det_rows_tc|desc_rows_tc|post_tc" -s 7 -c 7 -o /tmp/r02g_infer $CMD > gpurun_out/r02g_ncu_full.log 2>&1
echo "infer full rc=$?"; tail -2 gpurun_out/r02g_ncu_full.log
ncu -i /tmp/r02g_infer.ncu-rep --page raw --csv > gpurun_out/r02g_infer_raw.csv 2>/dev/null
export F3D_TRAIN_GRAPH=0
TCMD="python tools/train_bench.py"
timeout 120 $TCMD > gpurun_out/r02g_train_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:"lin_tc|wgrad_tc_kernel|bn_bwd_apply|bn_bwd_reduce_kernel|bn_apply" -s 47 -c 47 -o /tmp/r02g_train $TCMD > gpurun_out/r02g_ncu_train.log 2>&1
echo "train full rc=$?"; tail -2 gpurun_out/r02g_ncu_train.log
ncu -i /tmp/r02g_train.ncu-rep --page raw --csv > gpurun_out/r02g_train_raw.csv 2>/dev/null
ls -la gpurun_out/r02g_*raw.csv
