"""Launch the four ViT-B GEMM shapes a few times (for ncu): python tools/gemm_shapes.py [iters] [cta_pair] [block_n] [images]"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops
from peekvit_b200._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16, PK_EPI_BIAS_RESID_F32
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pair = int(sys.argv[2]) if len(sys.argv) > 2 else 2
bn = int(sys.argv[3]) if len(sys.argv) > 3 else 256
M = 197 * (int(sys.argv[4]) if len(sys.argv) > 4 else 256)
torch.manual_seed(0)
for (N, K, epi, name) in [(2304, 768, PK_EPI_BIAS_BF16, "qkv"), (768, 768, PK_EPI_BIAS_RESID_F32, "proj"),
                          (3072, 768, PK_EPI_BIAS_GELU_BF16, "fc1"), (768, 3072, PK_EPI_BIAS_RESID_F32, "fc2")]:
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") * 0.1
    bf = epi in (PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if bf else torch.float32)
    for _ in range(iters):
        ops.gemm(a, w, bias, out, epi, resid=None if bf else out, block_n=bn, cta_pair=pair)
    torch.cuda.synchronize()
    print(name, "flag", ops.device_flag())
