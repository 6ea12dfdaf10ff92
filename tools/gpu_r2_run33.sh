#!/bin/bash
timeout 600 python -m pytest tests/test_sparse_kernels_gpu.py tests/test_finetune_gpu.py -m gpu -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_bf16x2_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q -k "moe" 2>&1 | tail -3
for g in 1 0; do
  echo "ONEPASS=$g"
  PK_MOE_ROUTE_ONEPASS=$g timeout 600 python tools/variants_bench.py --batch 2048 --steps 10 --skip rank,residual,avit 2>&1 | grep -v "^$" | tail -3
done
