#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sparse_kernels_gpu.py tests/test_models_gpu.py -m gpu -q --timeout 900 2>&1 | grep -v "^$" > gpurun_out/r2_run8_pytest.log; grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/r2_run8_pytest.log | head -20; grep -n "^E  " gpurun_out/r2_run8_pytest.log | head -30
for f in 1 0; do
  echo "MOE_FUSED_SCATTER=$f"
  PEEKVIT_B200_MOE_FUSED_SCATTER=$f timeout 600 python tools/variants_bench.py --batch 2048 --steps 10 --skip rank,residual,avit 2>&1 | grep -v "^$" | tail -4
done
