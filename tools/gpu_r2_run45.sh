#!/bin/bash
# validation: full GPU suite, smoke, bench (both arms)
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -i "smoke" | tail -12
timeout 900 python bench.py > gpurun_out/r2_run45_bench.json 2> gpurun_out/r2_run45_bench.err; echo "bench rc $?"; tail -c 600 gpurun_out/r2_run45_bench.err
