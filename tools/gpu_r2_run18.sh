#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_finetune_gpu.py -m gpu -q -s --timeout 600 2>&1 | grep -v "^$" | grep -n "passed\|failed\|FAILED\|rel err\|^E  " | head -60
