#!/bin/bash
mkdir -p gpurun_out
for fam in residual avit; do
  if [ $fam = residual ]; then CMD="python tools/residual_run.py 0.4 512"; else CMD="python tools/avit_run.py"; fi
  PEEKVIT_B200_CUDA_GRAPHS=0 timeout 300 $CMD > gpurun_out/r2_run28_${fam}_plain.log 2>&1 && \
  PEEKVIT_B200_CUDA_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_run28_launches_${fam}.csv $CMD > gpurun_out/r2_run28_ncu_${fam}.log 2>&1
  echo "ncu $fam rc=$?"; tail -2 gpurun_out/r2_run28_${fam}_plain.log
done
