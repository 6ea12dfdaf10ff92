#!/bin/bash
timeout 100 python tools/attn_tcq_one.py 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_tcq -s 3 -c 1 -o gpurun_out/r2_run60_tcq python tools/attn_tcq_one.py > gpurun_out/r2_run60_ncu.log 2>&1
tail -2 gpurun_out/r2_run60_ncu.log; ls -la gpurun_out/r2_run60_tcq.ncu-rep
