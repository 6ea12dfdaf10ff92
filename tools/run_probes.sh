#!/bin/bash
# Run every kernel probe in its own process under a timeout; logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for fam in "$@"; do
  echo "=== probe $fam ==="
  timeout 300 python tools/probe.py $fam --json gpurun_out/probe_$fam.json > gpurun_out/probe_$fam.log 2>&1
  echo "exit=$?" >> gpurun_out/probe_$fam.log
  tail -n 60 gpurun_out/probe_$fam.log
done
