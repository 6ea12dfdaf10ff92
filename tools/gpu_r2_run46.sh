#!/bin/bash
timeout 200 python tools/debug_calib.py 2>&1 | tail -14
timeout 300 python -m pytest tests/test_finetune_gpu.py -m gpu -q --timeout 200 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_run46_bench.json 2> gpurun_out/r2_run46_bench.err; echo "bench rc $?"; tail -c 300 gpurun_out/r2_run46_bench.err
