"""A few launches of the general (mma.sync) attention kernel (for ncu): python tools/attn_general_run.py [B H dh N]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops
B, H, dh, N = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (64, 8, 32, 785)
D = H * dh
qkv = torch.randn(B * N, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, out, B, H, dh, seq_len=N, impl=1)
torch.cuda.synchronize()
