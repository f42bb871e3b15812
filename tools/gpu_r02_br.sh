#!/bin/bash
# round 2, GPU session br: grouped windowed ball query (4 KB bitmap per warp, 24 warps per SM) -- op + NMS tests, W4 flow, kernel list with the
# grouped kernel and with the whole-cloud-bitmap kernel before it (F3D_BQ_WHOLE_CLOUD_BITMAPS=1), C5 file flow on one GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_nms_gpu.py -m gpu -x -q > gpurun_out/r02br_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/r02br_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python tools/w4_kitti.py > gpurun_out/r02br_w4.jsonl 2> gpurun_out/r02br_w4.err; echo "w4 rc=$?"; cut -c1-330 gpurun_out/r02br_w4.jsonl; tail -3 gpurun_out/r02br_w4.err
timeout 300 python tools/w4_profile.py 2>&1 | grep "f3d::" | cut -c1-60,140-200 | head -6
F3D_BQ_WHOLE_CLOUD_BITMAPS=1 timeout 300 python tools/w4_profile.py 2>&1 | grep "f3d::bq" | cut -c1-60,140-200 | head -3
timeout 300 python tools/c5_kitti_multi.py --scans 128 --check 4 > gpurun_out/r02br_c5_1gpu.json 2> gpurun_out/r02br_c5_1gpu.err; echo "c5 rc=$?"; cut -c1-700 gpurun_out/r02br_c5_1gpu.json; tail -2 gpurun_out/r02br_c5_1gpu.err
