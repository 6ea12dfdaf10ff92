"""Pipeline timeline of the ragged tcgen05 attention kernel (CTA 0, its first 16 units):
    PK_ATT_TRACE=1 python tools/attn_trace_tcr.py [uniform_len | ragged]
Events per warp role (cycles after the first stamp); see tcr_trace() in csrc/pk_attention_tc.cu for the slot layout."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PK_ATT_TRACE"] = "1"
from peekvit_b200 import ops, _lib
mode = sys.argv[1] if len(sys.argv) > 1 else "197"
H, dh = 12 if mode != "ragged" else 6, 64
D = H * dh
if mode == "ragged":
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(40, 125, (512,), generator=g).tolist()
    B, rows = len(lens), sum(lens)
    cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), device="cuda", dtype=torch.int32)
    kw = dict(cu_seqlens=cu, max_seq_len=199, key_mult=torch.ones(rows, device="cuda"),
              extra_kv=(torch.randn(2 * D, device="cuda") * 0.3).to(torch.bfloat16), extra_mult=torch.full((B,), 100.0, device="cuda"))
else:
    N = int(mode)
    B, rows = 256, 256 * N
    kw = dict(seq_len=N)
qkv = torch.randn(rows, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.zeros(rows, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, out, B, H, dh, impl=3, **kw)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.attention(qkv, out, B, H, dh, impl=3, **kw)
b.record()
torch.cuda.synchronize()
print("us per launch", a.elapsed_time(b) * 100)
buf = np.zeros(16 * 16 * 8, dtype=np.uint64)
_lib.check(_lib.load().pk_attention_trace(buf.ctypes.data), "trace")
t = buf.reshape(16, 16, 8).astype(np.int64)
t0 = t[t > 0].min()
t = np.where(t > 0, t - t0, -1)
roles = [(0, "tma  ", ["top", "qk_slot_free", "v_slot_free"]),
         (1, "mma0 ", ["top", "qk_ready", "s_free", "qk_issued", "p_ready", "v_ready"]),
         (2, "mma1 ", ["top", "qk_ready", "s_free", "qk_issued", "p_ready", "v_ready"]),
         (3, "patch", ["lm_built", "qk_full", "v_full"]),
         (4, "sm r0h0", ["top", "s_full", "max_own", "max_all", "p_written", "p2_first_ld", "p2_group0", "p2_loop_end"]),
         (5, "sm r0h1", ["top", "s_full", "max_own", "max_all", "p_written", "p2_first_ld", "p2_group0", "p2_loop_end"]),
         (6, "sm r1h0", ["top", "s_full", "max_own", "max_all", "p_written", "p2_first_ld", "p2_group0", "p2_loop_end"]),
         (7, "sm r1h1", ["top", "s_full", "max_own", "max_all", "p_written", "p2_first_ld", "p2_group0", "p2_loop_end"]),
         (8, "out q0", ["top", "o_full", "o_read", "stored"]),
         (11, "out q3", ["top", "o_full", "o_read", "stored"])]
for k in range(6, 12):
    print(f"--- unit {k}")
    for slot, name, evs in roles:
        vals = {n: int(t[k, slot, e]) for e, n in enumerate(evs) if t[k, slot, e] >= 0}
        if vals:
            print(f"  {name:8s}", vals)
