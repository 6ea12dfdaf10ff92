#!/bin/bash
# One GPU visit: tests, smoke, bench (+ reference arm), then -- only after the same command exited 0 plain -- the ncu passes.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
# launch list of one 256-image micro-batch forward (eager launches so every kernel is visible to ncu)
SMALL="env PEEKVIT_B200_CUDA_GRAPHS=0 python bench.py --steps 1 --warmup 3 --batch 512 --no-cpu-baseline"
$SMALL > gpurun_out/plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 220 -c 80 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
$SMALL > gpurun_out/plain_small2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_pair|attention_tc|layernorm" -s 10 -c 8 -o gpurun_out/prof_model $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | head -40
