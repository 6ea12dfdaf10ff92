"""Top-1 agreement and logit error of the CUDA path against the fp32 CPU oracle on N ViT-B/16 images."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import peekvit_oracle as po, weights as ow
from peekvit_b200.models import VisionTransformer
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
sd = ow.make_state_dict("vit", cfg, seed=4321)
images = ow.synthetic_images(N, 224, seed=1234)
t0 = time.time()
ref = torch.cat([po.forward("vit", sd, cfg, images[s:s + 32])[0] for s in range(0, N, 32)])
t_cpu = time.time() - t0
model = VisionTransformer(**cfg); model.load_state_dict(sd); model = model.cuda().eval()
top2 = ref.topk(2, dim=1).values
margin = (top2[:, 0] - top2[:, 1])
for mode in ("bf16", "fp32"):
    model.pk_precision = mode
    logits = model(images.cuda()).cpu()
    err = ((logits - ref).abs().max() / ref.abs().max()).item()
    agree = (logits.argmax(1) == ref.argmax(1)).float().mean().item()
    dis = (logits.argmax(1) != ref.argmax(1)).nonzero().flatten()
    print(json.dumps(dict(mode=mode, images=N, rel_err=err, top1_agreement=agree, disagreements=int(dis.numel()),
                          oracle_margin_at_disagreements=[round(float(margin[i]), 5) for i in dis[:8]],
                          median_margin=float(margin.median()), max_abs_logit=float(ref.abs().max()), cpu_s=round(t_cpu, 1))))
# context: what the reference's own bf16 path (.to(bfloat16)) gives against the same fp32 oracle on this GPU
sd16 = {k: v.cuda().bfloat16() for k, v in sd.items()}
with torch.no_grad():
    logits = torch.cat([po.vit_forward(sd16, cfg, images[s:s + 128].cuda().bfloat16())[0].float().cpu() for s in range(0, N, 128)])
err = ((logits - ref).abs().max() / ref.abs().max()).item()
dis = (logits.argmax(1) != ref.argmax(1)).nonzero().flatten()
print(json.dumps(dict(mode="reference torch ops, .to(bfloat16), eager on this GPU", images=N, rel_err=err,
                      top1_agreement=1 - dis.numel() / N, disagreements=int(dis.numel()),
                      oracle_margin_at_disagreements=[round(float(margin[i]), 5) for i in dis[:8]],
                      max_margin_at_disagreements=round(float(margin[dis].max()), 5) if dis.numel() else 0.0)))
