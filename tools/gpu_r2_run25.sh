#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sanitize_r2.py > gpurun_out/r2_run25_plain.log 2>&1; echo "plain rc=$?"; tail -4 gpurun_out/r2_run25_plain.log
timeout 1500 compute-sanitizer --tool memcheck python tools/sanitize_r2.py > gpurun_out/r2_run25_sanitizer.log 2>&1; echo "sanitizer rc=$?"; grep -c "ok\|finetune\|routed" gpurun_out/r2_run25_sanitizer.log; tail -6 gpurun_out/r2_run25_sanitizer.log
