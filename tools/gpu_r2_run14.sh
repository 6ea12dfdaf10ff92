#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q 2>&1 | tail -2
timeout 300 python tools/attn_ragged_bench.py --json gpurun_out/r2_run14_attn_ragged.json 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    name, js = line.split(' ', 1)
    try: d = json.loads(js)
    except Exception: print(line.strip()); continue
    print(name, 'len', round(d['mean_len']), *[f\"{k}={v['us']:.1f}us/{v['tflops']:.0f}TF\" for k, v in d.items() if isinstance(v, dict) and 'us' in v])
"
timeout 600 python bench.py --steps 10 --warmup 3 --no-variants --no-cpu-baseline --no-gpu-reference > gpurun_out/r2_run14_bench.json 2> gpurun_out/r2_run14_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_run14_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'clocks')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'])
PY
