#!/bin/bash
# per-sample split between the ragged kernels: kernel tests, the A-ViT config test (with / without the quad kernel), variants A/B
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "attention" 2>&1 | tail -4
timeout 200 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q --timeout 120 -k "avit" 2>&1 | tail -6
PK_ATT_TCQ=0 timeout 200 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q --timeout 120 -k "avit" 2>&1 | tail -6
timeout 400 python tools/variants_bench.py --batch 2048 --steps 10 --skip rank,moe 2>&1 | grep -v "^$" | cut -c1-110
PK_ATT_TCQ=0 timeout 400 python tools/variants_bench.py --batch 2048 --steps 10 --skip rank,moe 2>&1 | grep -v "^$" | cut -c1-110
