#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_run22_bench.json 2> gpurun_out/r2_run22_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2_run22_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_run22_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks', 'device_flag')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'gpu_eager', (d.get('gpu_eager_reference') or {}).get('value'), 'cpu', d['cpu_baseline']['value'])
print('modes', {k: round(v['value']) for k, v in (d.get('precision_modes') or {}).items()})
for k, v in (d.get('variants') or {}).items():
    if isinstance(v, dict):
        print(k, round(v.get('value', 0)), 'x_dense', v.get('x_dense'), 'clk', v.get('clocks', {}).get('sm_mhz'), v.get('error', ''))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_run22_bench_reference.json 2>> gpurun_out/r2_run22_bench.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r2_run22_bench_reference.json
