#!/bin/bash
# round 2, GPU session bt: grouped windowed ball query (power-of-two index windows; other windows keep the whole-cloud-bitmap kernel) -- op + NMS
# tests, the grouped kernel forced onto windows of 3072 / 6144 points (diagnosis), W4 flow, kernel list with both kernels, C5 on one GPU
# (F3D_BQ_GRP_ANY_WINDOW existed at that commit only)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_nms_gpu.py -m gpu -x -q > gpurun_out/r02bt_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/r02bt_pytest.log
echo "== any window:"; F3D_BQ_GRP_ANY_WINDOW=1 timeout 120 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "70000 and not 170000" 2>&1 | tail -1
F3D_BQ_GRP_ANY_WINDOW=1 timeout 120 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "170000" 2>&1 | tail -1
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python tools/w4_kitti.py > gpurun_out/r02bt_w4.jsonl 2> gpurun_out/r02bt_w4.err; echo "w4 rc=$?"; grep bf16x3 gpurun_out/r02bt_w4.jsonl | cut -c1-330; tail -3 gpurun_out/r02bt_w4.err
timeout 300 python tools/w4_profile.py 2>&1 | grep "f3d::" | cut -c1-60,140-200 | head -6
F3D_BQ_WHOLE_CLOUD_BITMAPS=1 timeout 300 python tools/w4_profile.py 2>&1 | grep "f3d::bq" | cut -c1-60,140-200 | head -3
timeout 300 python tools/c5_kitti_multi.py --scans 128 --check 4 > gpurun_out/r02bt_c5_1gpu.json 2> gpurun_out/r02bt_c5_1gpu.err; echo "c5 rc=$?"; cut -c1-400 gpurun_out/r02bt_c5_1gpu.json; tail -2 gpurun_out/r02bt_c5_1gpu.err
