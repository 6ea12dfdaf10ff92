#!/bin/bash
mkdir -p gpurun_out
for m in 197 ragged 50; do echo "=== $m"; timeout 120 python tools/attn_trace_tcr.py $m 2>&1 | tail -80; done > gpurun_out/r2_run9_trace.txt 2>&1
tail -5 gpurun_out/r2_run9_trace.txt
timeout 300 python -m pytest tests/test_reference_utils.py tests/test_kernels_gpu.py -m gpu -q -k "reference or ragged" 2>&1 | tail -3
