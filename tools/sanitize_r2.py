"""Round-2 kernels under compute-sanitizer: the dense-layout ResidualViT cases (row_scale_add, gate plan on unit multiplicities),
the device-routed ragged attention on both sides of the threshold, the MoE fc2 with the fused un-permute, one fine-tuning step
(every backward kernel, both attention-backward variants):  compute-sanitizer --tool memcheck python tools/sanitize_r2.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
from golden_cases import CASES, build_case
from oracle import weights as ow
from peekvit_b200 import ops
from peekvit_b200.finetune import FineTuner
from peekvit_b200.models import VisionTransformer, add_noise, build_model
dev = "cuda:0"
for name in ("residual_skip_attention", "residual_skip_mlp_fixed", "residual_skip_mlp_add_input", "residual_gumbel_modes",
             "residual_two_cls_cal05", "residual_noise_snr", "residual_noise_token_drop", "residual_learnable_cal04", "moevit"):
    case = CASES[name]
    sd, images = build_case(case)
    model = build_model("vitmoe" if case["family"] == "moevit" else "residualvit", case["cfg"])
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    if case.get("budget") is not None:
        model.set_budget(case["budget"])
    if case.get("noise"):
        add_noise(model, **case["noise"])
    for mode in ("bf16", "bf16x2"):
        model.pk_precision = mode
        out = model(images.to(dev))
        torch.cuda.synchronize()
        print(name, mode, "ok", float(out.abs().max()), "flag", ops.device_flag(), flush=True)
# ragged attention, routed on the device
H, dh = 6, 64
D = H * dh
lens = [150, 33, 198, 77, 120, 5]
rows = sum(lens)
qkv = torch.zeros(rows + 16, 3 * D, device=dev, dtype=torch.bfloat16)
qkv[:rows] = torch.randn(rows, 3 * D, device=dev).to(torch.bfloat16)
cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), device=dev, dtype=torch.int32)
km = torch.ones(rows + 16, device=dev)
ekv = torch.randn(2 * D, device=dev).to(torch.bfloat16)
em = torch.full((len(lens),), 3.0, device=dev)
rd = torch.tensor([rows], device=dev, dtype=torch.int32)
for thr in (1, 10 ** 9):
    out = torch.zeros(rows + 16, D, device=dev, dtype=torch.bfloat16)
    ops.attention(qkv, out, len(lens), H, dh, cu_seqlens=cu, max_seq_len=199, key_mult=km, extra_kv=ekv, extra_mult=em, route_rows=rd,
                  route_min_rows=thr)
    torch.cuda.synchronize()
    print("routed attention", thr, float(out.float().abs().max()), "flag", ops.device_flag(), flush=True)
# fine-tuning steps: tcgen05 forward + mma.sync attention backward (head_dim 64), CUDA-core attention backward (head_dim 32)
for cfg, B in ((dict(image_size=64, patch_size=8, num_layers=2, num_heads=2, hidden_dim=128, mlp_dim=256, num_classes=10, num_class_tokens=2), 5),
               (dict(image_size=32, patch_size=8, num_layers=2, num_heads=2, hidden_dim=64, mlp_dim=128, num_classes=10), 3),
               (dict(image_size=224, patch_size=16, num_layers=1, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000), 2)):
    m = VisionTransformer(**cfg)
    m.load_state_dict(ow.make_state_dict("vit", cfg, seed=1))
    m = m.to(dev).train()
    ft = FineTuner(m, micro_batch=4)
    loss, _ = ft.forward_backward(torch.randn(B, 3, cfg["image_size"], cfg["image_size"], device=dev), torch.zeros(B, dtype=torch.long, device=dev))
    torch.cuda.synchronize()
    print("finetune", cfg["hidden_dim"], float(loss), float(m.class_tokens.grad.abs().max()), "flag", ops.device_flag(), flush=True)
