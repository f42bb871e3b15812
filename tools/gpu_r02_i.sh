#!/bin/bash
# round 2, GPU session i (one 8 x B200 box): bench.py (inference + training step with the NCCL all-reduce) at 8 and 4 ranks, and C5 --
# KITTI-shape scans through the inference.py file flow, sharded over 8 ranks, byte-compared with a single-process pass
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; }
nvidia-smi topo -m > gpurun_out/r02i_topo.txt 2>&1
timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 20 --warmup 3" > gpurun_out/r02i_bench_8gpu.json 2> gpurun_out/r02i_bench_8gpu.err
echo "bench8 rc=$?"
timeout 600 bash -c "$(declare -f run); run 4 bench.py --gpus 4 --steps 20 --warmup 3" > gpurun_out/r02i_bench_4gpu.json 2> gpurun_out/r02i_bench_4gpu.err
echo "bench4 rc=$?"
timeout 600 bash -c "$(declare -f run); run 8 tools/c5_kitti_multi.py --scans 128 --check 6" > gpurun_out/r02i_c5_8gpu.json 2> gpurun_out/r02i_c5_8gpu.err
echo "c5 rc=$?"
timeout 300 python tools/c5_kitti_multi.py --scans 16 --check 2 > gpurun_out/r02i_c5_1gpu.json 2> gpurun_out/r02i_c5_1gpu.err
echo "c5-1 rc=$?"
cut -c1-400 gpurun_out/r02i_bench_8gpu.json; tail -3 gpurun_out/r02i_bench_8gpu.err; cut -c1-300 gpurun_out/r02i_bench_4gpu.json; cat gpurun_out/r02i_c5_8gpu.json | cut -c1-600; tail -3 gpurun_out/r02i_c5_8gpu.err; cut -c1-400 gpurun_out/r02i_c5_1gpu.json
