#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_sparse_kernels_gpu.py -m gpu -q 2>&1 | tail -2
timeout 300 python tools/attn_ragged_bench.py --json gpurun_out/r2_run27_attn_ragged.json 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    name, js = line.split(' ', 1)
    try: d = json.loads(js)
    except Exception: print(line.strip()); continue
    print(name, 'len', round(d['mean_len']), *[f\"{k}={v['us']:.1f}us/{v['tflops']:.0f}TF\" for k, v in d.items() if isinstance(v, dict) and 'us' in v])
"
timeout 600 python tools/variants_bench.py --batch 2048 --steps 10 --skip rank,moe 2>&1 | grep -v "^$" | cut -c1-120
