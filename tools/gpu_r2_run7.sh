#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | grep -v "^$" > gpurun_out/r2_run7_pytest.log; grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/r2_run7_pytest.log | head -40; grep -n "^E  " gpurun_out/r2_run7_pytest.log | head -60
