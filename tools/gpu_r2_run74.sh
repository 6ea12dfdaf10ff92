#!/bin/bash
timeout 300 python -m pytest tests/test_bf16x2_gpu.py -m gpu -q --timeout 250 -s -k "north_star" 2>&1 | grep -E "4096 images|^E |passed|failed|64 images" | head
timeout 300 python -m pytest tests/test_bf16x2_gpu.py -m gpu -q --timeout 250 -s -k "north_star" 2>&1 | grep -E "4096 images|^E |passed|failed" | head
