#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -5
timeout 400 python tools/variants_bench.py --batch 2048 --steps 10 --skip rank 2>&1 | grep -v "^$" | cut -c1-110
