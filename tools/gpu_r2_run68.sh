#!/bin/bash
timeout 900 python bench.py > gpurun_out/r2_run68_bench.json 2> gpurun_out/r2_run68_bench.err; echo "bench rc $?"; tail -c 200 gpurun_out/r2_run68_bench.err
