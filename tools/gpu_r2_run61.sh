#!/bin/bash
timeout 300 ncu --set full --clock-control none -k regex:attention_fwd -s 3 -c 1 -o gpurun_out/r2_run61_dh32 python tools/attn_dh32_one.py > gpurun_out/r2_run61_ncu.log 2>&1
tail -1 gpurun_out/r2_run61_ncu.log
