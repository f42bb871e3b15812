#!/bin/bash
# round 2, GPU session bq (one 8 x B200 box): C5 -- KITTI-shape scans through the inference.py file flow sharded over 8 ranks (byte-compared with a
# single-process pass) after the NMS / ball-query pass
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; }
timeout 300 bash -c "$(declare -f run); run 8 tools/c5_kitti_multi.py --scans 256 --check 4" > gpurun_out/r02bq_c5_8gpu.json 2> gpurun_out/r02bq_c5_8gpu.err
echo "c5 rc=$?"; cut -c1-700 gpurun_out/r02bq_c5_8gpu.json; tail -2 gpurun_out/r02bq_c5_8gpu.err
