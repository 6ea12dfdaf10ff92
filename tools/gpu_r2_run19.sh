#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_finetune_gpu.py -m gpu -q --timeout 600 2>&1 | tail -3
python tools/finetune_run.py 512 128 3 2>&1 | tail -1
python tools/finetune_run.py 512 256 3 2>&1 | tail -1
