#!/bin/bash
# full GPU suite + variants with the quad kernel routed in
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4
timeout 600 python tools/variants_bench.py --batch 2048 --steps 10 --skip rank,moe 2>&1 | grep -v "^$" | cut -c1-110
PK_ATT_TCQ=0 timeout 600 python tools/variants_bench.py --batch 2048 --steps 10 --skip rank,moe 2>&1 | grep -v "^$" | cut -c1-110
