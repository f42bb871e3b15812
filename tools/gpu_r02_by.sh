#!/bin/bash
# round 2, GPU session by (the round's last GPU-seconds): W4 flow and smoke at HEAD with the hoisted shuffle of the grouped ball query
mkdir -p gpurun_out
timeout 40 python tools/w4_kitti.py > gpurun_out/r02by_w4.jsonl 2> gpurun_out/r02by_w4.err; echo "w4 rc=$?"; grep bf16x3 gpurun_out/r02by_w4.jsonl | cut -c1-330
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02by_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02by_smoke.log
