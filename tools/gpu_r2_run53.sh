#!/bin/bash
for mb in 128 256 512; do timeout 200 python tools/finetune_run.py 512 $mb 3 2>&1 | tail -1; done
timeout 200 python tools/finetune_residual_run.py 1024 1024 3 2>&1 | tail -1
