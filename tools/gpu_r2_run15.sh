#!/bin/bash
mkdir -p gpurun_out
for d in 0 1; do echo "=== diag $d"; PK_TCR_DIAG=$d PK_ATT_TRACE=1 timeout 120 python tools/attn_trace_tcr.py 197 2>&1 | grep "us per\|sm r0h0\|mma0" | head -7; done
