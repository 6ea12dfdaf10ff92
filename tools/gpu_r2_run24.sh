#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_run24_bench_2gpu.json 2> gpurun_out/r2_run24_err.log; echo "bench rc=$?"; tail -3 gpurun_out/r2_run24_err.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_run24_bench_2gpu.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step')}, 'e2e', d['e2e']['value'])
for k, v in (d.get('variants') or {}).items():
    if isinstance(v, dict):
        print(k, round(v.get('value', 0)), v.get('error', ''))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>> gpurun_out/r2_run24_err.log | tail -c 300; echo "ref rc=$?"
