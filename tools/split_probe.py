"""fp32 GEMM emulated on the bf16 tcgen05 kernel with 3-way split operands: accuracy probe."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops
from peekvit_b200._lib import PK_EPI_BIAS_F32
torch.manual_seed(0)
def split3(x):
    h = x.to(torch.bfloat16); r = x - h.float()
    m = r.to(torch.bfloat16); r = r - m.float()
    l = r.to(torch.bfloat16)
    return h, m, l
for (M, N, K) in [(1024, 768, 768), (1024, 768, 3072), (4096, 2304, 768)]:
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / math.sqrt(K); b = torch.randn(N, device="cuda") * 0.1
    ref = (a.double() @ w.double().t() + b.double())
    ref32 = a @ w.t() + b
    ah, am, al = split3(a); wh, wm, wl = split3(w)
    out = torch.empty(M, N, device="cuda")
    res = {}
    for name, (pa, pw) in {"1 term (bf16)": ([ah], [wh]), "3 terms": ([ah, ah, am], [wh, wm, wh]),
                           "6 terms": ([ah, ah, am, ah, al, am], [wh, wm, wh, wl, wh, wm])}.items():
        A = torch.cat(pa, 1).contiguous(); W = torch.cat(pw, 1).contiguous()
        ops.gemm(A, W, b, out, PK_EPI_BIAS_F32)
        res[name] = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    # small terms first: the big h*h products are added last
    A = torch.cat([am, al, ah, am, ah, ah], 1).contiguous(); W = torch.cat([wm, wh, wl, wh, wm, wh], 1).contiguous()
    ops.gemm(A, W, b, out, PK_EPI_BIAS_F32)
    res["6 terms, small first"] = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    res["torch fp32 matmul"] = ((ref32.double() - ref).abs().max() / ref.abs().max()).item()
    print((M, N, K), {k: f"{v:.2e}" for k, v in res.items()}, "flag", ops.device_flag())
