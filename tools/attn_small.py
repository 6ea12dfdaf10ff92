import sys, math, torch
sys.path.insert(0, ".")
from peekvit_b200 import ops
def t(fn, it=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it * 1e3
for (B, H, N) in [(512, 12, 99), (512, 12, 80), (512, 12, 50), (512, 12, 33), (512, 12, 26), (512, 12, 128), (512, 6, 197)]:
    D = H * 64
    qkv = torch.randn(B * N, 3 * D, device="cuda").to(torch.bfloat16)
    o1 = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16); o2 = torch.zeros_like(o1)
    ops.attention(qkv, o1, B, H, 64, seq_len=N, impl=1); ops.attention(qkv, o2, B, H, 64, seq_len=N, impl=2)
    err = ((o1.float() - o2.float()).abs().max() / o1.float().abs().max()).item()
    print(f"B{B} H{H} N{N}: general {t(lambda: ops.attention(qkv, o1, B, H, 64, seq_len=N, impl=1)):.1f}us tc {t(lambda: ops.attention(qkv, o2, B, H, 64, seq_len=N, impl=2)):.1f}us diff {err:.2e} flag {ops.device_flag()}")
