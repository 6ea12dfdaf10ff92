"""One gate-regime fine-tuning step of ResidualViT-S (config C kwargs) for timing / ncu launch lists:
    python tools/finetune_residual_run.py [batch] [micro_batch] [steps] [reg] [nocal]
(reg=1: with the mask regulariser; nocal: skip the gate calibration -- hundreds of forwards -- when only a launch list is wanted)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from peekvit_b200 import ops
from peekvit_b200.finetune import FineTuner
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
reg = len(sys.argv) > 4 and sys.argv[4] == "1"
dev = torch.device("cuda", 0)
cfg = dict(bench.CFG_S, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
           add_input=False, residual_layers=["attention+mlp"] * 12)
m = bench.make_model("residualvit", cfg, dev)
x = torch.randn(B, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (B,), device=dev)
if "nocal" not in sys.argv:
    bench.calibrate_residual_gates_(m, 0.5, x[:32], target=0.5)
m.train()
ft = FineTuner(m, micro_batch=mb)
opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=1e-3)


def regulariser(model):      # utils/losses.py:111-142 solo_mse(per_layer=False, strict=False), weight 0.01 (crossentropy_mse.yaml)
    sp = torch.stack([blk.mask.mean(dim=(1, 2)) for blk in model.encoder.layers]).mean()
    b = model.current_budget
    return 0.01 * (torch.relu(sp - b) ** 2).sum().mul(2 - b).mean()


def step():
    opt.zero_grad()
    loss, _ = ft.forward_backward(x, y, extra_loss=regulariser if reg else None)
    opt.step()
    return loss
for _ in range(2):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    loss = step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
print(f"B {B} mb {mb} reg {int(reg)}: ms/step {ms:.1f} img/s {B / ms * 1e3:.0f} loss {loss.item():.4f} flag {ops.device_flag()} "
      f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
