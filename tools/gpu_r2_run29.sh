#!/bin/bash
for occ in 3 4 5; do
  echo "=== OCC $occ"
  PK_ATT_GENERAL_OCC=$occ timeout 300 python tools/attn_ragged_bench.py 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    name, js = line.split(' ', 1)
    try: d = json.loads(js)
    except Exception: continue
    if 'residual' in name or 'avit' in name: print(name, 'len', round(d['mean_len']), *[f\"{k}={v['us']:.1f}us\" for k, v in d.items() if isinstance(v, dict) and 'us' in v])
"
done
