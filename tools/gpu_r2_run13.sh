#!/bin/bash
mkdir -p gpurun_out
python tools/attn_run.py 3 3 256 > gpurun_out/r2_run13_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tcr -s 3 -c 1 -o gpurun_out/prof_tcr python tools/attn_run.py 3 3 256 > gpurun_out/r2_run13_ncu.log 2>&1
echo "rc=$?"; cat gpurun_out/r2_run13_plain.log; tail -3 gpurun_out/r2_run13_ncu.log
