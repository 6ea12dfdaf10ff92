#!/bin/bash
cd tools && timeout 120 python attn_occ32.py 2>&1 | tail -6
