"""General (mma.sync) and tcgen05 attention kernels against torch's scaled_dot_product_attention on the same shapes
(context only: SDPA is library code, not used by the product)."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops
def t(fn, it=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it * 1e3
for (B, H, dh, N) in [(64, 8, 32, 785), (512, 12, 64, 197), (512, 6, 64, 110), (512, 12, 64, 50), (512, 8, 48, 197), (512, 3, 64, 197)]:
    D = H * dh
    qkv = torch.randn(B * N, 3 * D, device="cuda").to(torch.bfloat16)
    out = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
    q, k, v = (x.reshape(B, N, H, dh).transpose(1, 2).contiguous() for x in qkv.view(B, N, 3 * D).split(D, dim=-1))
    us_gen = t(lambda: ops.attention(qkv, out, B, H, dh, seq_len=N, impl=1))
    us_auto = t(lambda: ops.attention(qkv, out, B, H, dh, seq_len=N))
    us_sdpa = t(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
    fl = 4.0 * B * H * N * N * dh
    print(f"B{B} H{H} dh{dh} N{N}: general {us_gen:.1f} us ({fl / us_gen / 1e6:.0f} TF/s) | dispatched {us_auto:.1f} us | torch SDPA {us_sdpa:.1f} us ({fl / us_sdpa / 1e6:.0f} TF/s)")
