"""What the reference's own code path would give on this B200: the oracle port (the reference's torch ops) run
with CUDA tensors in torch eager mode — fp32 (the reference as shipped), and bf16 via .to(bfloat16).  Context for
the headline number only; not part of bench.py.   python tools/torch_eager_baseline.py [batch]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import peekvit_oracle as po, weights as ow
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
sd32 = {k: v.cuda() for k, v in ow.make_state_dict("vit", cfg, seed=4321).items()}
x32 = torch.randn(B, 3, 224, 224, device="cuda")
res = {}
for name, sd, x in (("fp32", sd32, x32), ("bf16", {k: v.bfloat16() for k, v in sd32.items()}, x32.bfloat16()),
                    ("fp32_tf32_matmul", sd32, x32)):
    torch.backends.cuda.matmul.allow_tf32 = name == "fp32_tf32_matmul"
    torch.backends.cudnn.allow_tf32 = name == "fp32_tf32_matmul"
    with torch.no_grad():
        for _ in range(3): po.vit_forward(sd, cfg, x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): po.vit_forward(sd, cfg, x)
        b.record(); torch.cuda.synchronize()
    res[name] = round(5 * B / a.elapsed_time(b) * 1e3, 1)
print(json.dumps({"torch_eager_img_per_s": res, "batch": B, "torch": torch.__version__}))
