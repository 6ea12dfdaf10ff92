"""A few launches of the general attention kernel on config A's shape (64 x 8 heads x 785 tokens, head_dim 32) for ncu."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops
B, H, dh, N = 64, 8, 32, 785
D = H * dh
qkv = torch.randn(B * N, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
for _ in range(5):
    ops.attention(qkv, out, B, H, dh, seq_len=N, impl=1)
torch.cuda.synchronize()
print("flag", ops.device_flag())
