#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_run31_bench_8gpu.json 2> gpurun_out/r2_run31_err.log; echo "bench rc=$?"; tail -3 gpurun_out/r2_run31_err.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_run31_bench_8gpu.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step', 'clocks')}, 'e2e', d['e2e']['value'])
print('strong', d.get('strong_scaling'))
for k, v in (d.get('variants') or {}).items():
    if isinstance(v, dict):
        print(k, round(v.get('value', 0)), v.get('error', ''))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 tools/finetune_ddp_check.py > gpurun_out/r2_run31_finetune_ddp_8gpu.json 2>> gpurun_out/r2_run31_err.log; echo "ddp rc=$?"; cat gpurun_out/r2_run31_finetune_ddp_8gpu.json
