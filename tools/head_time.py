"""cls_head kernel timing (final LayerNorm + class-token sum + head) at the ViT-B/16 shape: python tools/head_time.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops
B, seq, D, C = 512, 197, 768, 1000
x = torch.randn(B * seq, D, device="cuda")
g, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
w, hb = torch.randn(C, D, device="cuda") * 0.02, torch.zeros(C, device="cuda")
out = torch.empty(B, C, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot = 0.0
for i in range(13):
    flush.zero_()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.cls_head(x, B, seq, 1, g, b, 1e-5, w, hb, out=out); e.record(); torch.cuda.synchronize()
    if i >= 3: tot += a.elapsed_time(e)
ref = torch.nn.functional.linear(torch.nn.functional.layer_norm(x.view(B, seq, D)[:, 0], (D,), g, b, 1e-5), w, hb)
print(f"PK_HEAD_CTAS_PER_SM={os.environ.get('PK_HEAD_CTAS_PER_SM', 'default')}: {tot / 10 * 1e3:.1f} us, err {float((out - ref).abs().max()):.2e}")
# the large-batch head: LayerNorm of the class rows -> split -> tensor-core GEMM
from peekvit_b200._lib import PK_EPI_BIAS_F32
w6 = ops.split3_weight(w)
feat = torch.empty(B, D, device="cuda"); f6 = torch.empty(B, 6 * D, device="cuda", dtype=torch.bfloat16); out2 = torch.empty(B, C, device="cuda")
def head_gemm():
    ops.cls_features(x, B, seq, 1, g, b, 1e-5, feat)
    ops.split3(feat, f6)
    ops.gemm(f6, w6, hb, out2, PK_EPI_BIAS_F32)
tot = 0.0
for i in range(13):
    flush.zero_()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); head_gemm(); e.record(); torch.cuda.synchronize()
    if i >= 3: tot += a.elapsed_time(e)
print(f"cls_features + split3 + GEMM: {tot / 10 * 1e3:.1f} us, err {float((out2 - ref).abs().max()):.2e}")
