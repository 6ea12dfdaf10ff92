"""fp32-accurate mode: error against the CPU oracle and throughput (ViT-B/16): python tools/fp32_mode.py [images]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import peekvit_oracle as po, weights as ow
from peekvit_b200.models import VisionTransformer
from peekvit_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
sd = ow.make_state_dict("vit", cfg, seed=4321)
model = VisionTransformer(**cfg); model.load_state_dict(sd); model = model.cuda().eval()
small = ow.synthetic_images(16, 224, seed=1234)
ref, _ = po.forward("vit", sd, cfg, small)
ref64, _ = po.forward("vit", {k: v.double() for k, v in sd.items()}, cfg, small.double())
for mode in ("bf16", "fp32"):
    model.pk_precision = mode
    out = model(small.cuda()).cpu()
    print(mode, "rel err vs fp32 oracle", f"{((out - ref).abs().max() / ref.abs().max()).item():.2e}",
          "vs fp64 oracle", f"{((out.double() - ref64).abs().max() / ref64.abs().max()).item():.2e}")
print("fp32 oracle vs fp64 oracle", f"{((ref.double() - ref64).abs().max() / ref64.abs().max()).item():.2e}")
x = torch.randn(N, 3, 224, 224, device="cuda")
for mode in ("bf16", "fp32"):
    model.pk_precision = mode
    for _ in range(2): model(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); model(x); model(x); b.record(); torch.cuda.synchronize()
    print(mode, "img/s", 2 * N / a.elapsed_time(b) * 1e3, "flag", ops.device_flag())
