"""Pipeline timeline of the quad-region ragged tcgen05 attention kernel (CTA 0, its units 8 .. 23):
    PK_ATT_TRACE=1 python tools/attn_trace_tcq.py"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PK_ATT_TRACE"] = "1"
from peekvit_b200 import ops, _lib
H, dh = 6, 64
D = H * dh
g = torch.Generator().manual_seed(0)
lens = torch.randint(40, 125, (512,), generator=g).tolist()
B, rows = len(lens), sum(lens)
cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), device="cuda", dtype=torch.int32)
kw = dict(cu_seqlens=cu, max_seq_len=max(lens), key_mult=torch.ones(rows, device="cuda"),
          extra_kv=(torch.randn(2 * D, device="cuda") * 0.3).to(torch.bfloat16), extra_mult=torch.full((B,), 100.0, device="cuda"))
qkv = torch.randn(rows, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.zeros(rows, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, out, B, H, dh, impl=4, **kw)
torch.cuda.synchronize()
buf = np.zeros(16 * 16 * 8, dtype=np.uint64)
_lib.check(_lib.load().pk_attention_trace(buf.ctypes.data), "trace")
t = buf.reshape(16, 16, 8).astype(np.int64)
t0 = t[t > 0].min()
t = np.where(t > 0, t - t0, -1)
roles = [(0, "qk tma ", ["top", "slot_free"]), (6, "v tma  ", ["top", "slot_free"]),
         (5, "qk patch", ["top", "lm_free", "lm_built", "qk_full"]),
         (1, "mma0", ["top", "qk_ready", "s_free", "qk_issued", "p_ready", "v_ready", "pv_issued"]),
         (2, "mma1", ["top", "qk_ready", "s_free", "qk_issued", "p_ready", "v_ready", "pv_issued"]),
         (3, "mma2", ["top", "qk_ready", "s_free", "qk_issued", "p_ready", "v_ready", "pv_issued"]),
         (4, "mma3", ["top", "qk_ready", "s_free", "qk_issued", "p_ready", "v_ready", "pv_issued"]),
         (8, "sm r0", ["top", "s_full", "max_done", "p_written"]), (9, "sm r1", ["top", "s_full", "max_done", "p_written"]),
         (10, "sm r2", ["top", "s_full", "max_done", "p_written"]), (11, "sm r3", ["top", "s_full", "max_done", "p_written"]),
         (12, "out q0", ["top", "o_full", "o_read", "stored"])]
for k in range(0, 8):
    print(f"--- unit {k + 8} (region {k & 3}) len {lens[(0 + (k + 8) * 148) // H] if False else ''}")
    for slot, name, evs in roles:
        vals = {n: int(t[k, slot, e]) for e, n in enumerate(evs) if t[k, slot, e] >= 0}
        if vals:
            print(f"  {name:8s}", vals)
