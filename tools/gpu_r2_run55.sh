#!/bin/bash
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_run55_launches_avit_s_mb2048.csv python tools/avit_run.py 2048 > gpurun_out/r2_run55_ncu.log 2>&1
tail -1 gpurun_out/r2_run55_ncu.log
