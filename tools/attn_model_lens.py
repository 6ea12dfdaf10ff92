"""Ragged attention on the REALISED per-sample row counts of ResidualViT-S (budget 0.2 / 0.4, 512 images): the per-sample split
(impl 0: quad-region tcgen05 kernel for <= 128 keys + the general kernel for the rest) against the general kernel alone.

    python tools/attn_model_lens.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import weights as ow  # noqa: E402
from peekvit_b200 import ops  # noqa: E402
from attn_ragged_bench import timed  # noqa: E402
from variants_bench import VITS, make  # noqa: E402

DEV = "cuda:0"


def bench(lens, H=6):
    dh, D = 64, H * 64
    B, rows = len(lens), sum(lens)
    g = torch.Generator(device=DEV).manual_seed(1)
    qkv = torch.randn(rows + 256, 3 * D, device=DEV, generator=g).to(torch.bfloat16)
    out = torch.zeros(rows + 256, D, device=DEV, dtype=torch.bfloat16)
    cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), device=DEV, dtype=torch.int32)
    km = torch.ones(rows + 256, device=DEV)
    km[cu[1:].long() - 1] = 37.0
    ekv = (torch.randn(2 * D, device=DEV, generator=g) * 0.3).to(torch.bfloat16)
    em = torch.full((B,), 20.0, device=DEV)
    tot = torch.tensor([rows], device=DEV, dtype=torch.int32)
    kw = dict(cu_seqlens=cu, max_seq_len=199, key_mult=km, extra_kv=ekv, extra_mult=em)
    r = {}
    r["split_impl0"] = timed(lambda: ops.attention(qkv, out, B, H, dh, route_rows=tot, route_min_rows=B * 140, **kw))
    r["general_impl1"] = timed(lambda: ops.attention(qkv, out, B, H, dh, impl=1, **kw))
    short = [n if n + 1 <= 128 else 0 for n in lens]
    cus = torch.tensor([0] + torch.tensor(short).cumsum(0).tolist(), device=DEV, dtype=torch.int32)
    r["quad_short_only"] = timed(lambda: ops.attention(qkv, out, B, H, dh, impl=4, **dict(kw, cu_seqlens=cus, max_seq_len=127)))
    r["general_short_only"] = timed(lambda: ops.attention(qkv, out, B, H, dh, impl=1, **dict(kw, cu_seqlens=cus)))
    return r


def main():
    g = torch.Generator(device=DEV).manual_seed(1234)
    images = torch.randn(512, 3, 224, 224, device=DEV, generator=g)
    cfg = dict(VITS, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
               residual_layers=["attention+mlp"] * 12)
    sd0 = ow.make_state_dict("residualvit", cfg, seed=4321)
    for budget in (0.2, 0.4):
        sd = ow.calibrate_residual_gates(sd0, cfg, budget)
        m = make("residualvit", cfg, sd)
        m.set_budget(budget)
        m(images)
        for li, blk in enumerate(m.encoder.layers):
            mask = getattr(blk, "mask", None)
            if mask is None or li not in (1, 4, 8, 11):
                continue
            kept = (mask > 0).float().sum(dim=(1, 2)).long().cpu()
            lens = (kept + 3).tolist()            # class + budget token, kept image tokens, the ghost row
            t = torch.tensor(lens).float()
            st = dict(mean=float(t.mean()), min=int(t.min()), max=int(t.max()), frac_long=float((t + 1 > 128).float().mean()),
                      p10=float(t.quantile(0.1)), p90=float(t.quantile(0.9)))
            print(f"budget {budget} layer {li}", json.dumps(st), json.dumps({k: round(v, 1) for k, v in bench(lens).items()}), flush=True)
    print("flag", ops.device_flag())


if __name__ == "__main__":
    main()
