#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "quad or three_way" --timeout 60 2>&1 | tail -12
timeout 120 python tools/attn_ragged_bench.py --json gpurun_out/r2_run34_attn_ragged.json 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    name, js = line.split(' ', 1)
    try: d = json.loads(js)
    except Exception: print(line.strip()); continue
    print(name, 'len', round(d['mean_len']), *[f\"{k}={v['us']:.1f}us/{v['tflops']:.0f}TF\" if 'us' in v else f\"{k}=ERR {v}\" for k, v in d.items() if isinstance(v, dict)], 'flag', d.get('device_flag'))
"
