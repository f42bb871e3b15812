#!/bin/bash
# one ncu --set full capture of the first instance(s) of every hot-path kernel of the bench step + the launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
CMD="python bench.py --steps 2 --warmup 1 --graph 0 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$1.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fps_group|bq_grid|gather_point|det_rows_tc|desc_rows_tc|post_tc" -s 8 -c 8 -o gpurun_out/prof_$1 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
