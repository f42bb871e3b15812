#!/bin/bash
# round 2, GPU session bl: NMS top-K by radix select; nms_keep: float32 look first, warps walk the union of their candidates through a shared tile; bbox + max in one launch -- NMS tests, W4 flow, kernel list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nms_gpu.py -m gpu -x -q > gpurun_out/r02bl_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -15 gpurun_out/r02bl_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python tools/w4_kitti.py > gpurun_out/r02bl_w4.jsonl 2> gpurun_out/r02bl_w4.err; echo "w4 rc=$?"; cat gpurun_out/r02bl_w4.jsonl; tail -3 gpurun_out/r02bl_w4.err
timeout 300 python tools/w4_profile.py 2>&1 | grep "f3d::" | cut -c1-60,140-200 | head -14
