#!/bin/bash
timeout 200 python tools/residual_run.py 0.4 2048 2>&1 | tail -2
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_run54_launches_residualvit_s_b04_mb2048.csv python tools/residual_run.py 0.4 2048 > gpurun_out/r2_run54_ncu.log 2>&1
tail -1 gpurun_out/r2_run54_ncu.log
