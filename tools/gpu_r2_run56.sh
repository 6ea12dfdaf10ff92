#!/bin/bash
for f in 1 2 0; do echo "LN_FOLD=$f"; PEEKVIT_B200_LN_FOLD=$f timeout 100 python tools/vits_run.py 2048 8 2>&1 | tail -1; done
