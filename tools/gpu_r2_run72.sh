#!/bin/bash
timeout 300 python -m pytest tests/test_models_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q --timeout 200 -k "evaluate or eval" 2>&1 | tail -3
