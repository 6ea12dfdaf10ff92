"""Diagnostic: attention-MoE fixture — routing agreement with the reference and logit error per sample."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
from golden_cases import CASES, build_case
from peekvit_b200.models import build_model
from peekvit_b200 import runner
for seed in (20, 21, 22, 23):
    case = dict(CASES["moevit_attn"], weight_seed=seed)
    sd, images = build_case(case)
    from oracle import peekvit_oracle as po
    ref, oaux = po.forward("moevit", sd, case["cfg"], images)
    model = build_model("vitmoe", case["cfg"]); model.load_state_dict(sd); model = model.cuda().eval()
    out = model(images.cuda()).cpu()
    err = (out - ref).abs().amax(1) / ref.abs().max()
    agree = {}
    for i, blk in enumerate(model.encoder.layers):
        for nm, moe, key in (("attn", blk.self_attention, "attn_gating"), ("mlp", blk.mlp, "mlp_gating")):
            if moe.num_experts > 1:
                agree[f"{nm}{i}"] = float((moe.gating_probs.argmax(-1).cpu() == oaux[key][i].argmax(-1)).float().mean())
    print(seed, "max|ref|", float(ref.abs().max()), "err/sample", [f"{e:.4f}" for e in err.tolist()], agree)
