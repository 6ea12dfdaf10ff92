"""Error budget of the operand formats (CPU simulation, no GPU): which roundings of a ViT-B/16 forward produce the logits
error / top-1 flips of the bf16 mode, and what a 2-term split or fp16 operands buy.

    python tools/precision_budget.py [--images 256]

Every GEMM accumulates in fp32 (as the TMEM accumulator does); only the *operands* are rounded:
  lin  = A and W of the five linear GEMMs (patch, in-proj, out-proj, fc1, fc2)
  att  = q, k, v and the softmax probabilities p of the attention core
Formats: f32 (no rounding), bf16, fp16, bf16x2 (hi + lo, both bf16: 16 significant bits).
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import weights as ow  # noqa: E402

CFG = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)


def rnd(fmt):
    if fmt == "f32":
        return lambda x: x
    if fmt == "bf16":
        return lambda x: x.bfloat16().float()
    if fmt == "fp16":
        return lambda x: x.half().float()
    if fmt == "bf16x2":
        def f(x):
            h = x.bfloat16().float()
            return h + (x - h).bfloat16().float()
        return f
    raise ValueError(fmt)


def forward(sd, images, lin="f32", att="f32", sites=None):
    """sites: optional dict site -> fmt overriding lin/att for one site (qkv, out, fc1, fc2, patch, att)."""
    sites = sites or {}
    r = {s: rnd(sites.get(s, lin)) for s in ("patch", "qkv", "out", "fc1", "fc2")}
    ra = rnd(sites.get("att", att))
    D, H = CFG["hidden_dim"], CFG["num_heads"]
    dh = D // H
    p = CFG["patch_size"]
    B = images.shape[0]
    patches = F.unfold(images, p, stride=p).transpose(1, 2)                         # [B, P, 3pp]
    x = F.linear(r["patch"](patches), r["patch"](sd["conv_proj.weight"].reshape(D, -1)), sd["conv_proj.bias"])
    x = torch.cat([sd["class_tokens"].expand(B, -1, -1), x], 1) + sd["encoder.pos_embedding"]
    for i in range(CFG["num_layers"]):
        lp = f"encoder.layers.{i}"
        ap = lp + ".self_attention.self_attention"
        a = F.layer_norm(x, (D,), sd[lp + ".ln_1.weight"], sd[lp + ".ln_1.bias"], 1e-5)
        qkv = F.linear(r["qkv"](a), r["qkv"](sd[ap + ".in_proj_weight"]), sd[ap + ".in_proj_bias"])
        q, k, v = [ra(t).reshape(B, -1, H, dh).transpose(1, 2) for t in qkv.split(D, -1)]
        s = (q @ k.transpose(-1, -2)) * dh ** -0.5
        pr = torch.softmax(s, -1)
        # the kernel rounds the un-normalised exp to the operand format and divides by the fp32 row sum afterwards
        m = s.amax(-1, keepdim=True)
        e = torch.exp(s - m)
        o = (ra(e) @ v) / e.sum(-1, keepdim=True)
        o = o.transpose(1, 2).reshape(B, -1, D)
        x = x + F.linear(r["out"](o), r["out"](sd[ap + ".out_proj.weight"]), sd[ap + ".out_proj.bias"])
        a = F.layer_norm(x, (D,), sd[lp + ".ln_2.weight"], sd[lp + ".ln_2.bias"], 1e-5)
        h = F.gelu(F.linear(r["fc1"](a), r["fc1"](sd[lp + ".mlp.fc1.weight"]), sd[lp + ".mlp.fc1.bias"]))
        x = x + F.linear(r["fc2"](h), r["fc2"](sd[lp + ".mlp.fc2.weight"]), sd[lp + ".mlp.fc2.bias"])
    c = F.layer_norm(x[:, 0], (D,), sd["encoder.ln.weight"], sd["encoder.ln.bias"], 1e-5)
    return F.linear(c, sd["head.weight"], sd["head.bias"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    sd = ow.make_state_dict("vit", CFG, seed=4321)
    images = ow.synthetic_images(args.images, 224, seed=77)
    runs = [("bf16 everywhere", dict(lin="bf16", att="bf16")),
            ("bf16 linear only", dict(lin="bf16", att="f32")),
            ("bf16 attention only", dict(lin="f32", att="bf16")),
            ("fp16 everywhere", dict(lin="fp16", att="fp16")),
            ("bf16x2 linear + bf16 attention", dict(lin="bf16x2", att="bf16")),
            ("bf16x2 linear + fp16 attention", dict(lin="bf16x2", att="fp16")),
            ("bf16x2 everywhere", dict(lin="bf16x2", att="bf16x2")),
            ("bf16, fc1+fc2 bf16x2", dict(lin="bf16", att="bf16", sites=dict(fc1="bf16x2", fc2="bf16x2"))),
            ("bf16, qkv+out bf16x2", dict(lin="bf16", att="bf16", sites=dict(qkv="bf16x2", out="bf16x2")))]
    with torch.no_grad():
        chunks = [images[s:s + 32] for s in range(0, args.images, 32)]
        ref = torch.cat([forward(sd, c) for c in chunks])
        scale = ref.abs().max().item()
        top2 = ref.topk(2, 1).values
        margin = (top2[:, 0] - top2[:, 1]) / scale
        print(f"reference: max|logit| {scale:.3f}, median top-2 margin {margin.median():.4f} of it", flush=True)
        res = {}
        for name, kw in runs:
            t0 = time.time()
            out = torch.cat([forward(sd, c, **kw) for c in chunks])
            err = ((out - ref).abs().max() / scale).item()
            rms = ((out - ref).pow(2).mean().sqrt() / scale).item()
            flips = int((out.argmax(1) != ref.argmax(1)).sum())
            res[name] = dict(max_rel_err=err, rms_rel_err=rms, top1_flips=flips, images=args.images)
            print(f"{name:36s} max {err:.3e} rms {rms:.3e} flips {flips}/{args.images}  ({time.time() - t0:.0f} s)", flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
