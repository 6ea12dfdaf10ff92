"""One fine-tuning step of ViT-B/16 (class tokens + head regime) for timing / ncu launch lists:
    python tools/finetune_run.py [batch] [micro_batch] [steps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from peekvit_b200 import ops
from peekvit_b200.finetune import FineTuner
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
m = bench.make_model("vit", bench.CFG_B, dev)
m.train()
ft = FineTuner(m, micro_batch=mb)
opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=1e-3)
x = torch.randn(B, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (B,), device=dev)
def step():
    opt.zero_grad()
    loss, _ = ft.forward_backward(x, y)
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
print(f"ms/step {ms:.1f} img/s {B / ms * 1e3:.0f} loss {loss.item():.4f} flag {ops.device_flag()} mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
