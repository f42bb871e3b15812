#!/bin/bash
# round 2, GPU session c: every parity figure the GPU tests produce (no -x), to set the tolerance table from data; the rest of the suite
mkdir -p gpurun_out
rm -f gpurun_out/r02c_parity_report.jsonl
F3D_PARITY_REPORT=gpurun_out/r02c_parity_report.jsonl timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02c_pytest_gpu.log
