#!/bin/bash
timeout 100 python tools/finetune_residual_run.py 256 256 1 0 nocal 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 900 --csv --log-file gpurun_out/r2_run44_launches_finetune_residual.csv python tools/finetune_residual_run.py 256 256 1 0 nocal > gpurun_out/r2_run44_ncu.log 2>&1
tail -1 gpurun_out/r2_run44_ncu.log
