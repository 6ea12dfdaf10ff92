#!/bin/bash
timeout 400 python -m pytest tests/test_finetune_gpu.py -m gpu -q --timeout 200 -s -k "residualvit_s_gate or rejects" 2>&1 | grep -E "flipped|^loss |passed|failed|^E  |worst" | head -40
