"""Time the four ViT-B GEMM shapes: python tools/gemm_time.py [images] [cta_pair] [block_n]"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops
from peekvit_b200._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16, PK_EPI_BIAS_RESID_F32
imgs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pair = int(sys.argv[2]) if len(sys.argv) > 2 else 2
bn = int(sys.argv[3]) if len(sys.argv) > 3 else 256
M = 197 * imgs
torch.manual_seed(0)
def time_ms(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
res = []
for (N, K, epi, name) in [(2304, 768, PK_EPI_BIAS_BF16, "qkv"), (768, 768, PK_EPI_BIAS_RESID_F32, "proj"),
                          (3072, 768, PK_EPI_BIAS_GELU_BF16, "fc1"), (768, 3072, PK_EPI_BIAS_RESID_F32, "fc2")]:
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") * 0.1
    bf = epi in (PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if bf else torch.float32)
    ms = time_ms(lambda: ops.gemm(a, w, bias, out, epi, resid=None if bf else out, block_n=bn, cta_pair=pair))
    res.append(f"{name} {ms*1e3:.1f}us {2.0*M*N*K/ms/1e9:.0f}TF")
print(f"imgs={imgs} pair={pair} bn={bn} env={os.environ.get('PK_GEMM_L2_PREFETCH','-')}:", " | ".join(res), "flag", ops.device_flag())
