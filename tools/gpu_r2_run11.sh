#!/bin/bash
mkdir -p gpurun_out
for m in 197 ragged; do echo "=== $m"; PK_ATT_TRACE=1 timeout 120 python tools/attn_trace_tcr.py $m 2>&1 | sed -n 1,40p; done > gpurun_out/r2_run11_trace.txt 2>&1
grep "sm r0h0\|sm r1h0\|us per" gpurun_out/r2_run11_trace.txt | head -14
