"""Pipeline timeline of the tcgen05 attention kernel (CTA 0): PK_ATT_TRACE=1 python tools/attn_trace.py"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PK_ATT_TRACE"] = "1"
from peekvit_b200 import ops, _lib
B, H, N, dh = 256, 12, 197, 64
D = H * dh
qkv = torch.randn(B * N, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(qkv, out, B, H, dh, seq_len=N, impl=2)
torch.cuda.synchronize()
buf = np.zeros(16 * 12 * 8, dtype=np.uint64)
_lib.check(_lib.load().pk_attention_trace(buf.ctypes.data), "trace")
t = buf.reshape(16, 12, 8).astype(np.int64)
t0 = t[t > 0].min()
t = np.where(t > 0, t - t0, -1)
names_mma = ["kv_full", "qk0_issued", "qk1_issued", "p0_ready", "p1_ready", "pv0_issued", "pv1_issued"]
names_sm = ["loop_top", "s_full", "max_done", "p_written", "o_full", "o_read", "stored", "staged"]
for it in range(5, 8):
    print(f"--- item {it}")
    print("  mma  ", {n: int(t[it, 1, e]) for e, n in enumerate(names_mma)})
    for w in (4, 5, 8, 10):
        print(f"  warp{w}", {n: int(t[it, w, e]) for e, n in enumerate(names_sm)})
