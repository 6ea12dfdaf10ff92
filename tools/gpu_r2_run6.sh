#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_ragged_bench.py --json gpurun_out/r2_run6_attn_ragged.json 2>&1 | tail -14
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | grep -v "^$" > gpurun_out/r2_run6_pytest.log; tail -12 gpurun_out/r2_run6_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_run6_bench.json 2> gpurun_out/r2_run6_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_run6_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_run6_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks', 'device_flag')})
    print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'gpu_eager', (d.get('gpu_eager_reference') or {}).get('value'))
    print('modes', d.get('precision_modes'))
    for k, v in (d.get('variants') or {}).items():
        if isinstance(v, dict):
            print(k, round(v['value']), 'x_dense', v.get('x_dense'), 'clk', v.get('clocks', {}).get('sm_mhz'))
except Exception as e:
    print('parse failed', e)
PY
# launch lists (eager launches) of the two ragged families with the new attention kernel
for fam in residual avit; do
  if [ $fam = residual ]; then CMD="python tools/residual_run.py 0.4 512"; else CMD="python tools/avit_run.py"; fi
  PEEKVIT_B200_CUDA_GRAPHS=0 timeout 300 $CMD > gpurun_out/r2_run6_${fam}_plain.log 2>&1 && \
  PEEKVIT_B200_CUDA_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_run6_launches_${fam}.csv $CMD > gpurun_out/r2_run6_ncu_${fam}.log 2>&1
  echo "ncu $fam rc=$?"; tail -2 gpurun_out/r2_run6_${fam}_plain.log
done
