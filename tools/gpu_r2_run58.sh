#!/bin/bash
timeout 200 python tools/gemm_gelu_ab.py 2>&1 | tail -4
