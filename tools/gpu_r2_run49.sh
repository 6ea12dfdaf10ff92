#!/bin/bash
timeout 100 python tools/vits_run.py 512 10 2>&1 | tail -1
timeout 100 python tools/vits_run.py 2048 5 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 120 --csv --log-file gpurun_out/r2_run49_launches_vits_dense.csv python tools/vits_run.py 512 2 > gpurun_out/r2_run49_ncu.log 2>&1
tail -1 gpurun_out/r2_run49_ncu.log
