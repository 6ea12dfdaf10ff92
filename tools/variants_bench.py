"""Throughput of the budgeted variants against their dense baselines (BASELINE.json configs 2-4).

    python tools/variants_bench.py [--batch 1024] [--steps 3] [--json gpurun_out/variants.json]

Device-resident synthetic images, random-init weights (oracle.weights, seeded).  ResidualViT gate biases are
calibrated per budget with the CPU oracle so that the realised keep fraction is ~ the budget (SURVEY.md §7.3 H7:
with random gates the keep fraction is otherwise 0 or 1).  Reported: images/s, speed-up over the dense model of the
same shape, realised tokens per layer.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import weights as ow  # noqa: E402  (weights + gate calibration only; nothing on the timed path)
from peekvit_b200 import ops, runner  # noqa: E402
from peekvit_b200.models import build_model  # noqa: E402

DEV = torch.device("cuda", 0)
VITB = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
VITS = dict(image_size=224, patch_size=16, num_layers=12, num_heads=6, hidden_dim=384, mlp_dim=1536, num_classes=1000)


def timed(model, images, steps):
    for _ in range(2):
        out = model(images)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        out = model(images)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    return images.shape[0] / ms * 1e3, out


def make(name, cfg, sd):
    m = build_model(name, cfg)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--json", default=None)
    ap.add_argument("--skip", default="")
    args = ap.parse_args()
    B = args.batch
    g = torch.Generator(device=DEV).manual_seed(1234)
    images = torch.randn(B, 3, 224, 224, device=DEV, generator=g)
    res = {}

    def report(k, **kw):
        res[k] = kw
        print(k, json.dumps(kw), flush=True)

    # ---- dense baselines
    dense_b, _ = timed(make("vit", VITB, ow.make_state_dict("vit", VITB, seed=4321)), images, args.steps)
    report("vit_b_16_dense", img_s=dense_b)
    dense_s, _ = timed(make("vit", VITS, ow.make_state_dict("vit", VITS, seed=4321)), images, args.steps)
    report("vit_s_16_dense", img_s=dense_s)

    # ---- config 3: RankViT on the ViT-B shape, rank layers [3, 6, 9]
    if "rank" not in args.skip:
        cfg = dict(VITB, rankvit_layers=[3, 6, 9])
        m = make("RankVisionTransformer", cfg, ow.make_state_dict("rankvit", cfg, seed=4321))
        for budget in (1.0, 0.5, 0.4, 0.25):
            m.set_budget(budget)
            v, _ = timed(m, images, args.steps)
            aux = {}
            runner.run(m, images[:64], aux)
            report(f"rankvit_b_budget{budget}", img_s=v, speedup_vs_dense=v / dense_b, tokens_per_layer=aux.get("seq_lens"))

    # ---- config 2: ResidualViT, ViT-S shape, learnable budget token, all layers gated
    if "residual" not in args.skip:
        cfg = dict(VITS, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
                   residual_layers=["attention+mlp"] * 12)
        sd0 = ow.make_state_dict("residualvit", cfg, seed=4321)
        for budget in (0.2, 0.4, 0.8, 1.0):
            sd = ow.calibrate_residual_gates(sd0, cfg, min(budget, 0.97))     # budget 1.0: gates calibrated to keep ~97 %
            m = make("residualvit", cfg, sd)
            m.set_budget(budget)
            v, _ = timed(m, images, args.steps)
            keep = [float((blk.mask > 0).float().mean()) for blk in m.encoder.layers if getattr(blk, "mask", None) is not None]
            report(f"residualvit_s_budget{budget}", img_s=v, speedup_vs_dense=v / dense_s, keep_fraction_per_layer=[round(k, 3) for k in keep])

    # ---- config 4: A-ViT halting and MoE expert MLPs on the ViT-S shape
    if "avit" not in args.skip:
        cfg = dict(VITS, eps=0.01, gate_scale=1.0, gate_center=1.5)      # random-init stand-in for a trained halting gate
        m = make("adavit", cfg, ow.make_state_dict("adavit", cfg, seed=4321))
        v, _ = timed(m, images, args.steps)
        cnt = m.encoder.counter_token
        report("avit_s", img_s=v, speedup_vs_dense=v / dense_s, mean_layers_per_token=float(cnt.float().mean()) if cnt is not None else None)
    if "moe" not in args.skip:
        cfg = dict(VITS, mlp_moes=[4] * 12)
        m = make("vitmoe", cfg, ow.make_state_dict("moevit", cfg, seed=4321))
        v, _ = timed(m, images, args.steps)
        report("moevit_s_4experts", img_s=v, speedup_vs_dense=v / dense_s)
    report("device_flag", flag=ops.device_flag())
    if args.json:
        os.makedirs(os.path.dirname(args.json), exist_ok=True)
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
