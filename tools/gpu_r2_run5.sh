#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/sanitize_tcr.py 2>&1 | tail -5
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_tcr.py > gpurun_out/r2_run5_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -8 gpurun_out/r2_run5_sanitizer.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -k "ragged or attention" 2>&1 | grep -v "^$" > gpurun_out/r2_run5_pytest_attn.log; tail -15 gpurun_out/r2_run5_pytest_attn.log
