#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q --timeout 600 -s 2>&1 | grep -v "^$" > gpurun_out/r2_run2_pytest.log; tail -5 gpurun_out/r2_run2_pytest.log
