"""One ResidualViT-S forward at a budget (for ncu launch lists): python tools/residual_run.py [budget] [batch]
Weights and gate calibration as in bench.py's variants block (projection biases bisected on the device)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from peekvit_b200 import ops
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 0.4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
cfg = dict(bench.CFG_S, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
           add_input=False, residual_layers=["attention+mlp"] * 12)
m = bench.make_model("residualvit", cfg, dev)
x = torch.randn(B, 3, 224, 224, device=dev)
bench.calibrate_residual_gates_(m, budget, x[:32], target=min(budget, 0.97))
m.set_budget(budget)
for _ in range(3): out = m(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = m(x); b.record(); torch.cuda.synchronize()
keep = [round(float((blk.mask > 0).float().mean()), 2) for blk in m.encoder.layers]
print("keep", keep)
print("ms", a.elapsed_time(b), "img/s", B / a.elapsed_time(b) * 1e3, "flag", ops.device_flag())
# for `ncu --profile-from-start off`: one more forward inside a profiler range (the calibration above is ~30k launches)
torch.cuda.profiler.start()
out = m(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
