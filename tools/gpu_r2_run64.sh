#!/bin/bash
# final validation of the round: full GPU suite, smoke, bench
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -i "smoke" | tail -10
timeout 900 python bench.py > gpurun_out/r2_run64_bench.json 2> gpurun_out/r2_run64_bench.err; echo "bench rc $?"; tail -c 200 gpurun_out/r2_run64_bench.err
