#!/bin/bash
mkdir -p gpurun_out
python tools/finetune_run.py 128 128 2 > gpurun_out/r2_run21_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 420 --csv --log-file gpurun_out/r2_run21_launches_finetune.csv python tools/finetune_run.py 128 128 2 > gpurun_out/r2_run21_ncu.log 2>&1
echo rc=$?; cat gpurun_out/r2_run21_plain.log; tail -2 gpurun_out/r2_run21_ncu.log
