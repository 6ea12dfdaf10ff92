import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops
from attn_ragged_bench import timed
B, H, dh, N = 64, 8, 32, 785
D = H * dh
qkv = torch.randn(B * N, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
q, k, v = (x.reshape(B, N, H, dh).transpose(1, 2).contiguous() for x in qkv.view(B, N, 3 * D).split(D, dim=-1))
fl = 4.0 * B * H * N * N * dh
for occ in ("3", "4", "5", "6"):
    os.environ["PK_ATT_GENERAL_OCC32"] = occ
    us = timed(lambda: ops.attention(qkv, out, B, H, dh, seq_len=N, impl=1))
    print(f"occ {occ}: {us:.1f} us ({fl / us / 1e6:.0f} TF/s)")
us = timed(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
print(f"torch SDPA: {us:.1f} us ({fl / us / 1e6:.0f} TF/s)")
