#!/bin/bash
for rep in 1 2; do
for lib in lib_nokmv_tcq lib_kmv; do
  echo "=== $lib (rep $rep)"
  PEEKVIT_B200_LIB=$PWD/tools/probes/$lib.so timeout 300 python tools/attn_ragged_bench.py 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    name, js = line.split(' ', 1)
    try: d = json.loads(js)
    except Exception: continue
    if 'b0.4' in name or 'uniform_50' in name or '197' in name: print(name, *[f\"{k}={v['us']:.1f}\" for k, v in d.items() if isinstance(v, dict) and 'us' in v])
"
done
done
