#!/bin/bash
# round 2, GPU session bw: nms_keep with a branch-free float32 path -- NMS tests, NMS alone on a KITTI-shape scan
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nms_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python tools/nms_only.py 131072 50
timeout 200 python tools/w4_profile.py 2>&1 | grep "f3d::nms_keep" | cut -c1-60,140-200
