#!/bin/bash
for sp in "0.25,0.75" "0.125,0.875" "0.125,0.375,0.5" "0.0625,0.1875,0.75" "1.0"; do PEEKVIT_B200_HOST_FIRST_SPLIT=$sp timeout 100 python tools/e2e_small_batch.py 256 2>&1 | tail -1; done
