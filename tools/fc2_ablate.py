"""fc2 shape (M = 197*512, N = 768, K = 3072): in-place TMA-reduce variant vs LayerNorm-producer variant, with the
PK_GEMM_DEBUG ablation given in the environment (1 = no operand loads, 2 = no epilogue)."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops
from peekvit_b200._lib import PK_EPI_BIAS_RESID_F32, PK_EPI_BIAS_GELU_BF16
M, D, F = 197 * 512, 768, 3072
def time_us(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
x = torch.randn(M, D, device="cuda"); xb = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
stats = torch.zeros(M, ops.gemm_row_stat_parts(D), 2, device="cuda")
hid = (torch.randn(M, F, device="cuda") * 0.5).to(torch.bfloat16)
w2 = (torch.randn(D, F, device="cuda") / math.sqrt(F)).to(torch.bfloat16); b2 = torch.randn(D, device="cuda") * 0.1
a1 = (torch.randn(M, D, device="cuda")).to(torch.bfloat16)
w1 = (torch.randn(F, D, device="cuda") / math.sqrt(D)).to(torch.bfloat16); b1 = torch.randn(F, device="cuda") * 0.1
t_red = time_us(lambda: ops.gemm(hid, w2, b2, x, PK_EPI_BIAS_RESID_F32, resid=x))
t_prod = time_us(lambda: ops.gemm(hid, w2, b2, x, PK_EPI_BIAS_RESID_F32, resid=x, xb_out=xb, row_stats=stats))
t_fc1 = time_us(lambda: ops.gemm(a1, w1, b1, hid, PK_EPI_BIAS_GELU_BF16))
fl = 2.0 * M * D * F
print(f"debug={os.environ.get('PK_GEMM_DEBUG', '0')}: fc2 reduce {t_red:.1f} us ({fl / t_red / 1e6:.0f} TF/s) | fc2 producer {t_prod:.1f} us ({fl / t_prod / 1e6:.0f}) | fc1 gelu {t_fc1:.1f} us ({fl / t_fc1 / 1e6:.0f})")
