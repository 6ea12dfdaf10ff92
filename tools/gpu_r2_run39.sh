#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "per_sample_split" 2>&1 | tail -3
timeout 300 python tools/attn_model_lens.py 2>&1 | tail -12
