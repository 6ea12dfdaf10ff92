#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r2_run51_bench.json 2> gpurun_out/r2_run51_bench.err; echo "bench rc $?"; tail -c 300 gpurun_out/r2_run51_bench.err
