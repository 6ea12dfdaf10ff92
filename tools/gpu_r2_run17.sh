#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "ragged or attention" 2>&1 | tail -2
for thr in 140 0 100000; do
  echo "=== ATT_TCR_MIN_MEAN_ROWS=$thr"
  PEEKVIT_B200_ATT_TCR_MIN_MEAN_ROWS=$thr timeout 600 python tools/variants_bench.py --batch 2048 --steps 10 --skip rank,moe 2>&1 | grep -v "^$" | cut -c1-150
done
