#!/bin/bash
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_run48_bench_2gpu.json 2> gpurun_out/r2_run48_bench_2gpu.err; echo "rc $?"; tail -c 400 gpurun_out/r2_run48_bench_2gpu.err
