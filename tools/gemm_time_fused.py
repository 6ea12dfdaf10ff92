"""Time the LayerNorm-fused GEMM variants on the ViT-B shapes: python tools/gemm_time_fused.py [images]"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops
from peekvit_b200._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16, PK_EPI_BIAS_RESID_F32
imgs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M, D, F = 197 * imgs, 768, 3072
torch.manual_seed(0)
def time_ms(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
P = ops.gemm_row_stat_parts(D)
x = torch.randn(M, D, device="cuda")
xb = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
stats = torch.zeros(M, P, 2, device="cuda")
ops.row_stats_cast(x, xb, stats)
res = []
for (N, K, epi, name, mode) in [(2304, 768, PK_EPI_BIAS_BF16, "qkv", "cons"), (768, 768, PK_EPI_BIAS_RESID_F32, "proj", "prod"),
                                (3072, 768, PK_EPI_BIAS_GELU_BF16, "fc1", "cons"), (768, 3072, PK_EPI_BIAS_RESID_F32, "fc2", "prod")]:
    a = xb if K == D and mode == "cons" else (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") * 0.1
    c1 = torch.randn(N, device="cuda")
    if mode == "cons":
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms0 = time_ms(lambda: ops.gemm(a, w, bias, out, epi))
        ms1 = time_ms(lambda: ops.gemm(a, w, bias, out, epi, ln_stats=stats, ln_c1=c1, ln_dim=D, ln_eps=1e-5))
    else:
        ms0 = time_ms(lambda: ops.gemm(a, w, bias, x, epi, resid=x))
        ms1 = time_ms(lambda: ops.gemm(a, w, bias, x, epi, resid=x, xb_out=xb, row_stats=stats))
    res.append(f"{name} plain {ms0*1e3:.1f}us fused {ms1*1e3:.1f}us")
ln = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
g = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
ms_ln = time_ms(lambda: ops.layernorm(x, g, b, 1e-5, ln))
print(" | ".join(res), f"| layernorm {ms_ln*1e3:.1f}us", "flag", ops.device_flag())
