#!/bin/bash
# round 2, GPU session af: ncu evidence for the training step as it stands at the end of the round -- launch list of one eager step and one
# --set full capture of its large kernels (second eager step).  Raw CSVs come back; the summary is made by tools/ncu_summary.py.
mkdir -p gpurun_out
export F3D_TRAIN_GRAPH=0
TCMD="python tools/train_bench.py"
timeout 120 $TCMD > gpurun_out/r02af_train_plain.log 2>&1
echo "plain rc=$?"; tail -1 gpurun_out/r02af_train_plain.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02af_launches_train.csv $TCMD > gpurun_out/r02af_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"lin_tc|wgrad_tc_kernel|bn_bwd_apply|bn_bwd_reduce_kernel|bn_apply|conv3_fwd|pool_from_extremes|local_frames_kernel|loss_min" -s 50 -c 50 -o /tmp/r02af_train $TCMD > gpurun_out/r02af_ncu_train.log 2>&1
echo "train full rc=$?"; tail -2 gpurun_out/r02af_ncu_train.log
ncu -i /tmp/r02af_train.ncu-rep --page raw --csv > gpurun_out/r02af_train_raw.csv 2>/dev/null
ls -la gpurun_out/r02af_*
