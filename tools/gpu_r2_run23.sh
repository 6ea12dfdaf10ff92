#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --batch 512 --no-variants --no-cpu-baseline --no-gpu-reference"
PEEKVIT_B200_CUDA_GRAPHS=0 $CMD > gpurun_out/r2_run23_plain.log 2>&1 && \
PEEKVIT_B200_CUDA_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r2_run23_ncu_launches.csv $CMD > gpurun_out/r2_run23_ncu1.log 2>&1
echo "launch list rc=$?"
PEEKVIT_B200_CUDA_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_pair_kernel|attention_tc3|layernorm_bf16" -s 60 -c 12 -o gpurun_out/prof_r2_model $CMD > gpurun_out/r2_run23_ncu2.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/r2_run23_ncu2.log; ls -la gpurun_out/prof_r2_model.ncu-rep
