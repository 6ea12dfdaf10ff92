#!/bin/bash
timeout 300 python -m pytest tests/test_models_gpu.py -m gpu -q --timeout 200 -k "forward_host or uint8 or host" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_run69_bench.json 2> gpurun_out/r2_run69_bench.err; echo "bench rc $?"; tail -c 300 gpurun_out/r2_run69_bench.err
