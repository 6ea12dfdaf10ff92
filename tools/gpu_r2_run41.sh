#!/bin/bash
timeout 300 python -m pytest tests/test_models_gpu.py -m gpu -q --timeout 200 -s -k "short_sequence" 2>&1 | grep -v "^$" | tail -8
