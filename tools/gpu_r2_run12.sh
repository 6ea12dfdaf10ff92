#!/bin/bash
PK_ATT_TRACE=1 timeout 120 python tools/attn_trace.py 2>&1 | tail -24
