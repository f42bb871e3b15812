#!/bin/bash
# round 2, GPU session bv: kernel table of one KITTI-shape scan through the W4 flow at HEAD + ncu --set full of the grouped ball query in it
mkdir -p gpurun_out
timeout 200 python tools/w4_profile.py > gpurun_out/r02bv_w4_profile.txt 2>&1; echo "profile rc=$?"; grep "f3d::" gpurun_out/r02bv_w4_profile.txt | cut -c1-60,140-200 | head -8
timeout 120 python tools/w4_once.py 3 > gpurun_out/r02bv_plain.log 2>&1; rc=$?; tail -1 gpurun_out/r02bv_plain.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bq_grid_query_grp -s 6 -c 2 -o gpurun_out/r02bv_bq_grp python tools/w4_once.py 2 > gpurun_out/r02bv_ncu.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/r02bv_ncu.log
