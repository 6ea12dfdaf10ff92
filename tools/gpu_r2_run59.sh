#!/bin/bash
PK_ATT_SPLIT=1 timeout 400 python tools/variants_bench.py --batch 2048 --steps 10 --skip rank,moe 2>&1 | grep -E "residualvit_s_budget0.4|residualvit_s_budget0.8|avit" | cut -c1-90
timeout 400 python tools/variants_bench.py --batch 2048 --steps 10 --skip rank,moe 2>&1 | grep -E "residualvit_s_budget0.4|residualvit_s_budget0.8|avit" | cut -c1-90
