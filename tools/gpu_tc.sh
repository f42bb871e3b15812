#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -q --tb=short -s -k selftest > gpurun_out/tc_selftest.log 2>&1
echo "rc=$?" >> gpurun_out/tc_selftest.log
timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -q --tb=short -s -k detector > gpurun_out/tc_det.log 2>&1
echo "rc=$?" >> gpurun_out/tc_det.log
timeout 300 python bench.py --steps 10 --warmup 3 --precision bf16x3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "rc=$?" >> gpurun_out/bench_tc.err
tail -15 gpurun_out/tc_selftest.log; tail -25 gpurun_out/tc_det.log; cat gpurun_out/bench_tc.json; tail -3 gpurun_out/bench_tc.err
