"""Launch the tcgen05 attention a few times (for ncu): python tools/attn_run.py [iters] [impl] [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
H, N, dh = 12, 197, 64
D = H * dh
torch.manual_seed(0)
qkv = torch.randn(B * N, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
for _ in range(iters):
    ops.attention(qkv, out, B, H, dh, seq_len=N, impl=impl)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    ops.attention(qkv, out, B, H, dh, seq_len=N, impl=impl)
b.record(); torch.cuda.synchronize()
print("us", a.elapsed_time(b) / iters * 1e3, "flag", ops.device_flag())
