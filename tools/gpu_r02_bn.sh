#!/bin/bash
# round 2, GPU session bn: ncu --set full of the NMS kernels on one KITTI-shape scan (after the command has run plain)
mkdir -p gpurun_out
timeout 120 python tools/nms_only.py 131072 20 > gpurun_out/r02bn_nms_plain.log 2>&1; rc=$?; cat gpurun_out/r02bn_nms_plain.log | tail -2
if [ $rc -ne 0 ]; then exit 0; fi
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nms_ -s 11 -c 11 -o gpurun_out/r02bn_nms python tools/nms_only.py 131072 1 > gpurun_out/r02bn_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02bn_ncu.log; ls -la gpurun_out/r02bn_nms.ncu-rep
