#!/bin/bash
# ncu --set full (with source/SASS sampling) of one FPS launch at C3 for a given variant
mkdir -p gpurun_out
export TAG=${1:-cur}
timeout 120 python tools/fps_ab.py > gpurun_out/fps_plain.log 2>&1 || { tail -5 gpurun_out/fps_plain.log; exit 1; }
head -3 gpurun_out/fps_plain.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"fps_" -s 2 -c 1 -f -o gpurun_out/fps_$TAG python tools/fps_ab.py > gpurun_out/ncu_fps.log 2>&1
tail -2 gpurun_out/ncu_fps.log; ls -la gpurun_out/*.ncu-rep
