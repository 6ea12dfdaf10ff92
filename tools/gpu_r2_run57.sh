#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "gelu_epilogue_accuracy_pair" 2>&1 | grep -E "^E|Error" | head -20
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from peekvit_b200 import ops
from peekvit_b200._lib import PK_EPI_BIAS_GELU_BF16
M=1024
xs = torch.linspace(-10, 10, M * 64, device="cuda").to(torch.bfloat16).view(M, 64)
eye = torch.eye(64, device="cuda", dtype=torch.bfloat16)
out = torch.empty(M, 64, device="cuda", dtype=torch.bfloat16)
ops.gemm(xs, eye, None, out, PK_EPI_BIAS_GELU_BF16, cta_pair=2)
ref = torch.nn.functional.gelu(xs.float())
err = (out.float()-ref).abs()
bad = err > ref.abs()*2.0**-8 + 8e-6
print("bad", int(bad.sum()), "worst abs", float(err.max()), "at x", float(xs.flatten()[err.argmax()]), "ref", float(ref.flatten()[err.argmax()]), "out", float(out.flatten()[err.argmax()].float()))
idx = bad.flatten().nonzero().flatten()[:10]
for i in idx: print(float(xs.flatten()[i]), float(ref.flatten()[i]), float(out.flatten()[i].float()))
print("zero", float(out[xs==0].abs().max()), "big equal", bool((out[xs>9]==xs[xs>9]).all()))
PY
