#!/bin/bash
timeout 300 python -m pytest tests/test_finetune_gpu.py -m gpu -q --timeout 200 2>&1 | tail -2
timeout 200 python tools/finetune_run.py 512 256 3 2>&1 | tail -1
timeout 200 python tools/finetune_residual_run.py 512 512 3 2>&1 | tail -1
