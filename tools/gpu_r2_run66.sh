#!/bin/bash
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_run66_launches_moevit_s_mb2048.csv python tools/moe_run.py 2048 > gpurun_out/r2_run66_ncu.log 2>&1
tail -1 gpurun_out/r2_run66_ncu.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_run66_launches_vits_mb2048.csv python tools/vits_prof.py 2048 > gpurun_out/r2_run66_ncu2.log 2>&1
tail -1 gpurun_out/r2_run66_ncu2.log
