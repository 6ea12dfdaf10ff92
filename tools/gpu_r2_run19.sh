#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_finetune_gpu.py -m gpu -q -s --timeout 600 2>&1 | grep -n "passed\|failed\|FAILED\|rel err\|^E  " | head -40
python tools/finetune_run.py 512 128 3 2>&1 | tail -1
PK_ATT_BWD_MMA=0 python tools/finetune_run.py 512 128 3 2>&1 | tail -1
