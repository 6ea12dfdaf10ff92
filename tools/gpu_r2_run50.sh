#!/bin/bash
timeout 600 python tools/mb_sweep_vits.py 1024 2048 4096 --rank --4096 --only=residualvit_s_b0.4 --only=avit --only=rankvit_b_b0.5 2>&1 | grep -v "^$" | tail -5
