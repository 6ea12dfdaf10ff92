#!/bin/bash
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_run52_bench_reference.json 2> gpurun_out/r2_run52_ref.err; echo "ref rc $?"; tail -c 300 gpurun_out/r2_run52_bench_reference.json
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -i "smoke" | tail -3
