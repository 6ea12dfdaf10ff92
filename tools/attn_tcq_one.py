"""A few launches of the quad-region ragged attention kernel on the 40 - 125-row case (for `ncu --set full -k regex:attention_tcq`)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops
DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
lens = torch.randint(40, 125, (512,), generator=g).tolist()
H, dh = 6, 64
D = H * dh
rows = sum(lens)
gg = torch.Generator(device=DEV).manual_seed(1)
qkv = torch.randn(rows + 128, 3 * D, device=DEV, generator=gg).to(torch.bfloat16)
out = torch.zeros(rows + 128, D, device=DEV, dtype=torch.bfloat16)
cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), device=DEV, dtype=torch.int32)
km = torch.ones(rows + 128, device=DEV)
km[cu[1:].long() - 1] = 37.0
ekv = (torch.randn(2 * D, device=DEV, generator=gg) * 0.3).to(torch.bfloat16)
em = torch.full((512,), 100.0, device=DEV)
for _ in range(5):
    ops.attention(qkv, out, 512, H, dh, cu_seqlens=cu, max_seq_len=max(lens), key_mult=km, extra_kv=ekv, extra_mult=em, impl=4)
torch.cuda.synchronize()
print("flag", ops.device_flag(), "rows", rows)
