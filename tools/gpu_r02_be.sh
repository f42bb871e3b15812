#!/bin/bash
# round 2, GPU session be: ncu evidence after the last kernel pass -- launch list + one --set full capture of every hot inference kernel
# (serial eager step).  Raw CSV comes back; the summary is made by tools/ncu_summary.py.
mkdir -p gpurun_out
CMD="python bench.py --workload infer --steps 2 --warmup 1 --graph 0 --pipelined 0 --no-cpu-baseline"
$CMD > gpurun_out/r02be_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02be_launches_infer.csv $CMD > gpurun_out/r02be_ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r02be_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fps_group|bq_grid|det_rows_tc|desc_rows_tc|post_tc" -s 7 -c 7 -o /tmp/r02be_infer $CMD > gpurun_out/r02be_ncu_full.log 2>&1
echo "infer full rc=$?"; tail -2 gpurun_out/r02be_ncu_full.log
ncu -i /tmp/r02be_infer.ncu-rep --page raw --csv > gpurun_out/r02be_infer_raw.csv 2>/dev/null
ls -la gpurun_out/r02be_*raw.csv
