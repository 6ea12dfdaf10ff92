"""One MoE-ViT-S forward (for ncu launch lists): python tools/moe_run.py [batch]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PEEKVIT_B200_CUDA_GRAPHS", "0")
from oracle import weights as ow
from peekvit_b200.models import build_model
from peekvit_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = dict(image_size=224, patch_size=16, num_layers=12, num_heads=6, hidden_dim=384, mlp_dim=1536, num_classes=1000, mlp_moes=[4] * 12)
m = build_model("vitmoe", cfg); m.load_state_dict(ow.make_state_dict("moevit", cfg, seed=4321), strict=True); m = m.cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda")
for _ in range(3): out = m(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = m(x); b.record(); torch.cuda.synchronize()
print("ms", a.elapsed_time(b), "img/s", B / a.elapsed_time(b) * 1e3, "flag", ops.device_flag())
# for `ncu --profile-from-start off`: one more forward inside a profiler range
torch.cuda.profiler.start()
out = m(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
