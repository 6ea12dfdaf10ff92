#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/finetune_ddp_check.py > gpurun_out/r2_run20_finetune_ddp_2gpu.json 2> gpurun_out/r2_run20_err.log; echo "rc=$?"; cat gpurun_out/r2_run20_finetune_ddp_2gpu.json; tail -3 gpurun_out/r2_run20_err.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-variants --no-gpu-reference > gpurun_out/r2_run20_bench_2gpu.json 2>> gpurun_out/r2_run20_err.log; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_run20_bench_2gpu.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step', 'clocks')}, 'e2e', d['e2e']['value'], 'strong', d.get('strong_scaling'))
PY
