"""Dense ViT-S/16 forwards (for ncu launch lists / timing): python tools/vits_run.py [batch] [iters]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import weights as ow
from peekvit_b200.models import build_model
from peekvit_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = dict(image_size=224, patch_size=16, num_layers=12, num_heads=6, hidden_dim=384, mlp_dim=1536, num_classes=1000)
m = build_model("vit", cfg); m.load_state_dict(ow.make_state_dict("vit", cfg, seed=4321), strict=True); m = m.cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda")
for _ in range(3): out = m(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters): out = m(x)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
print("ms", ms, "img/s", B / ms * 1e3, "flag", ops.device_flag())
