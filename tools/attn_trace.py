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
buf = np.zeros(16 * 16 * 8, dtype=np.uint64)
_lib.check(_lib.load().pk_attention_trace(buf.ctypes.data), "trace")
t = buf.reshape(16, 16, 8).astype(np.int64)
t0 = t[t > 0].min()
t = np.where(t > 0, t - t0, -1)
names_mma = ["ready", "qk_issued", "p_ready", "pv_issued", "kv_full"]
names_sm = ["loop_top", "s_full", "max_done", "p_written", "max_own"]
names_out = ["r0_o_full", "r0_staged", "r0_stored", "r1_o_full", "r1_staged", "r1_stored"]
# column-split kernel (default): warps 4-7 = region 0 / column half 0, 8-11 = region 0 / half 1; output warps in slots 12-15
for it in range(7, 9):
    print(f"--- item {it}")
    print("  tma   ", {"qk_slot_free": int(t[it, 0, 0]), "v_slot_free": int(t[it, 0, 1])})
    for w in (1, 2):
        print(f"  mma{w - 1} ", {n: int(t[it, w, e]) for e, n in enumerate(names_mma)})
    for w in (4, 8):
        print(f"  soft{w}", {n: int(t[it, w, e]) for e, n in enumerate(names_sm)})
    for w in (12, 13, 14, 15):
        print(f"  out{w}", {n: int(t[it, w, e]) for e, n in enumerate(names_out)})
