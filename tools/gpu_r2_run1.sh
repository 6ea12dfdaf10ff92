#!/bin/bash
# Round 2, GPU visit 1: the whole GPU test suite (new BASELINE-shape parity tests included), smoke, the new bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_run1_pytest.log; tail -25 gpurun_out/r2_run1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_run1_bench.json 2> gpurun_out/r2_run1_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_run1_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_run1_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks', 'device_flag')})
    print('e2e', d['e2e'])
    print('roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'share_of_step')})
    print('gpu_eager_reference', d.get('gpu_eager_reference'))
    print('cpu_baseline', d.get('cpu_baseline'))
    print('modes', d.get('precision_modes'))
    for k, v in (d.get('variants') or {}).items():
        print(k, v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk not in ('note', 'gates')})
except Exception as e:
    print('parse failed', e)
PY
