#!/bin/bash
timeout 100 python -m pytest tests/test_finetune_gpu.py -m gpu -q --timeout 90 -k "reference_fixture" 2>&1 | tail -3
