#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/finetune_ddp_check.py residualvit > gpurun_out/r2_run67_finetune_ddp_gates_2gpu.json 2> gpurun_out/r2_run67.err; echo "rc $?"; cat gpurun_out/r2_run67_finetune_ddp_gates_2gpu.json | tail -1 | cut -c1-600; tail -3 gpurun_out/r2_run67.err | cut -c1-300
