"""One ResidualViT-S forward at a budget (for ncu launch lists): python tools/residual_run.py [budget] [batch] [family]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import weights as ow
from peekvit_b200.models import build_model
from peekvit_b200 import ops
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 0.4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
VITS = dict(image_size=224, patch_size=16, num_layers=12, num_heads=6, hidden_dim=384, mlp_dim=1536, num_classes=1000)
cfg = dict(VITS, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
           residual_layers=["attention+mlp"] * 12)
sd = ow.calibrate_residual_gates(ow.make_state_dict("residualvit", cfg, seed=4321), cfg, budget)
m = build_model("residualvit", cfg); m.load_state_dict(sd, strict=True); m = m.cuda().eval(); m.set_budget(budget)
x = torch.randn(B, 3, 224, 224, device="cuda")
for _ in range(3): out = m(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = m(x); b.record(); torch.cuda.synchronize()
print("ms", a.elapsed_time(b), "img/s", B / a.elapsed_time(b) * 1e3, "flag", ops.device_flag())
