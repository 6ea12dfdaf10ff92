"""A few small launches of the ragged tcgen05 attention kernel for compute-sanitizer: python tools/sanitize_tcr.py"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops  # noqa: E402

DEV = "cuda:0"
H, dh = 2, 64
D = H * dh
for lens, extra in (([3, 70, 130, 1], True), ([40, 17], False), ([200], True)):
    rows = sum(lens)
    qkv = torch.zeros(rows + 8, 3 * D, device=DEV, dtype=torch.bfloat16)
    qkv[:rows] = torch.randn(rows, 3 * D, device=DEV).to(torch.bfloat16)
    cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), device=DEV, dtype=torch.int32)
    km = torch.randint(1, 5, (rows + 8,), device=DEV).float()
    ekv = torch.randn(2 * D, device=DEV).to(torch.bfloat16) if extra else None
    em = torch.tensor([float(i % 3) for i in range(len(lens))], device=DEV) if extra else None
    out = torch.zeros(rows + 8, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, len(lens), H, dh, cu_seqlens=cu, max_seq_len=max(lens), key_mult=km, extra_kv=ekv, extra_mult=em, impl=3)
    torch.cuda.synchronize()
    out1 = torch.zeros_like(out)
    ops.attention(qkv, out1, len(lens), H, dh, cu_seqlens=cu, max_seq_len=max(lens), key_mult=km, extra_kv=ekv, extra_mult=em, impl=1)
    torch.cuda.synchronize()
    err = ((out.float() - out1.float()).abs().max() / out1.float().abs().max()).item()
    print(lens, "extra" if extra else "", "rel diff vs general kernel", err, "flag", ops.device_flag(), flush=True)
