#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "per_sample_split" 2>&1 | grep -E "^E|^FAILED|passed|failed|^tests.*Error" | head -60
