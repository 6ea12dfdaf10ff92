#!/bin/bash
timeout 200 python tools/finetune_residual_run.py 512 256 3 2>&1 | tail -2
timeout 200 python tools/finetune_residual_run.py 512 512 3 1 2>&1 | tail -2
timeout 200 python tools/finetune_residual_run.py 512 128 3 2>&1 | tail -2
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 3000 --csv --log-file gpurun_out/r2_run43_launches_finetune_residual.csv python tools/finetune_residual_run.py 256 256 1 > gpurun_out/r2_run43_ncu.log 2>&1
tail -2 gpurun_out/r2_run43_ncu.log
