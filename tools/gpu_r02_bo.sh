#!/bin/bash
# round 2, GPU session bo: ncu --set full of the windowed ball query inside the W4 flow of one KITTI-shape scan (after the command has run plain)
mkdir -p gpurun_out
timeout 120 python tools/w4_once.py 3 > gpurun_out/r02bo_plain.log 2>&1; rc=$?; tail -1 gpurun_out/r02bo_plain.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bq_grid_query_win -s 6 -c 2 -o gpurun_out/r02bo_bq python tools/w4_once.py 2 > gpurun_out/r02bo_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02bo_ncu.log; ls -la gpurun_out/r02bo_bq.ncu-rep
