"""Small forwards of every family and mode for compute-sanitizer: python tools/sanitize_small.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
from golden_cases import CASES, build_case
from peekvit_b200.models import build_model, add_noise
from peekvit_b200 import ops
NAMES = {"vit": "vit", "rankvit": "RankVisionTransformer", "residualvit": "residualvit", "adavit": "adavit", "moevit": "vitmoe",
         "eeresidualvit": "eeResidualvit"}
for name, case in CASES.items():
    sd, images = build_case(case)
    model = build_model(NAMES[case["family"]], case["cfg"])
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    if case.get("budget") is not None:
        model.set_budget(case["budget"])
    if case.get("noise"):
        add_noise(model, **case["noise"])
    for mode in ("bf16", "fp32"):
        model.pk_precision = mode
        out = model(images.cuda())
        out = out[-1] if isinstance(out, list) else out
        torch.cuda.synchronize()
        print(name, mode, "ok", float(out.abs().max()), "flag", ops.device_flag(), flush=True)
# one ViT-B/16-shaped layer pair so the CTA-pair GEMM, the tcgen05 attention and the token-row patch embedding run too
cfg = dict(image_size=224, patch_size=16, num_layers=2, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
from oracle import weights as ow
m = build_model("vit", cfg); m.load_state_dict(ow.make_state_dict("vit", cfg, seed=1)); m = m.cuda().eval()
print("vit_b16_2layers", float(m(torch.randn(4, 3, 224, 224, device="cuda")).abs().max()), "flag", ops.device_flag())
