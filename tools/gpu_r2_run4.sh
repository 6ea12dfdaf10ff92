#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_bf16x2_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q --timeout 900 -s 2>&1 | grep -v "^$" > gpurun_out/r2_run4_pytest.log; grep -n "config \|  bf16\|bf16x2 \|4096 images\|passed\|failed\|^E  " gpurun_out/r2_run4_pytest.log | head -80
timeout 600 python bench.py --steps 5 --warmup 3 --no-variants --no-cpu-baseline --no-gpu-reference > gpurun_out/r2_run4_bench.json 2> gpurun_out/r2_run4_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r2_run4_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_run4_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'device_flag')}, d['e2e']['value'])
    print('modes', d.get('precision_modes'))
except Exception as e:
    print('parse failed', e)
PY
