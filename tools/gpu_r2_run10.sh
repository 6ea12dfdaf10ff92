#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_sparse_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -3
timeout 300 python tools/attn_ragged_bench.py --json gpurun_out/r2_run10_attn_ragged.json 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    name, js = line.split(' ', 1)
    try: d = json.loads(js)
    except Exception: print(line.strip()); continue
    print(name, 'len', round(d['mean_len']), *[f\"{k}={v['us']:.1f}us/{v['tflops']:.0f}TF\" for k, v in d.items() if isinstance(v, dict) and 'us' in v])
"
for m in 197 ragged; do echo "=== $m"; PK_ATT_TRACE=1 timeout 120 python tools/attn_trace_tcr.py $m 2>&1 | sed -n 1,40p; done > gpurun_out/r2_run10_trace.txt 2>&1
