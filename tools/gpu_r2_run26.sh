#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_finetune_gpu.py tests/test_models_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q -s --timeout 600 2>&1 | grep -n "passed\|failed\|FAILED\|rankvit .*rel err\|^E  " | head -30
