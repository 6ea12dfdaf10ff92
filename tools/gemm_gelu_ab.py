"""fc1 shapes with and without the GELU in the epilogue (CUDA-graph timed): what the exact-erf epilogue costs on top of the
plain bias + bf16 epilogue of the same GEMM.  python tools/gemm_gelu_ab.py"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from peekvit_b200 import ops
from peekvit_b200._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16
from attn_ragged_bench import timed
for name, imgs, N, K in (("vit_b fc1", 512, 3072, 768), ("vit_s fc1", 512, 1536, 384), ("vit_s fc1 2048 imgs", 2048, 1536, 384)):
    M = 197 * imgs
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") * 0.1
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    r = {}
    for label, epi in (("bias", PK_EPI_BIAS_BF16), ("gelu", PK_EPI_BIAS_GELU_BF16)):
        us = timed(lambda: ops.gemm(a, w, bias, out, epi, cta_pair=2), n=10)
        r[label] = f"{us:.1f} us ({2.0 * M * N * K / us / 1e6:.0f} TF/s)"
    print(name, r, "flag", ops.device_flag(), flush=True)
