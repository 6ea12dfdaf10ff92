"""Micro-batch sweep on the ViT-S shaped models (dense, ResidualViT at two budgets, A-ViT, MoE; --vitb / --rank add the ViT-B
shaped dense model and RankViT): images/s at 2048 (or --4096) images per step for model.pk_micro_batch in the sizes given.

    python tools/mb_sweep_vits.py 512 1024 2048 [--vitb] [--rank] [--4096] [--only=<substring> ...]

The question was whether ViT-S blocks (HBM-bound at 512 images: the activations of one micro-batch exceed the 126 MB L2) gain
from L2-resident micro-batches.  They do not: smaller is slower everywhere, and the families that run on compacted / routed rows
gain from LARGER micro-batches (fewer launches, longer kernels) -- runner.SPARSE_MICRO_BATCH.  One fresh model per size, so a
workspace never holds more than one micro-batch size."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import weights as ow  # noqa: E402
from variants_bench import VITS, VITB, make, timed  # noqa: E402

DEV = torch.device("cuda", 0)
NIMG = 4096 if "--4096" in sys.argv else 2048
images = torch.randn(NIMG, 3, 224, 224, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1234))
def builders():
    yield "vit_s_dense", lambda: make("vit", VITS, ow.make_state_dict("vit", VITS, seed=4321))
    cfg = dict(VITS, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
               residual_layers=["attention+mlp"] * 12)
    sd0 = ow.make_state_dict("residualvit", cfg, seed=4321)
    for b in (0.4, 0.8):
        def mk(b=b):
            m = make("residualvit", cfg, ow.calibrate_residual_gates(sd0, cfg, b))
            m.set_budget(b)
            return m
        yield f"residualvit_s_b{b}", mk
    cfga = dict(VITS, eps=0.01, gate_scale=1.0, gate_center=1.5)
    yield "avit_s", lambda: make("adavit", cfga, ow.make_state_dict("adavit", cfga, seed=4321))
    cfgm = dict(VITS, mlp_moes=[4] * 12)
    yield "moevit_s", lambda: make("vitmoe", cfgm, ow.make_state_dict("moevit", cfgm, seed=4321))
    if "--vitb" in sys.argv:
        yield "vit_b_dense", lambda: make("vit", VITB, ow.make_state_dict("vit", VITB, seed=4321))
    if "--rank" in sys.argv:
        cfgr = dict(VITB, rankvit_layers=[3, 6, 9])
        for b in (0.5, 0.25):
            def mkr(b=b):
                m = make("RankVisionTransformer", cfgr, ow.make_state_dict("rankvit", cfgr, seed=4321))
                m.set_budget(b)
                return m
            yield f"rankvit_b_b{b}", mkr


only = [a[7:] for a in sys.argv if a.startswith("--only=")]
sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [64, 128, 192, 256, 384, 512, 1024]
for name, mk in builders():
    if only and not any(o in name for o in only):
        continue
    row = {}
    for mb in sizes:
        m = mk()                      # a fresh model per size: its workspace holds one micro-batch size only
        m.pk_micro_batch = mb
        v, _ = timed(m, images, 6)
        row[mb] = round(v)
        del m
        torch.cuda.empty_cache()
    print(name, json.dumps(row), f"peak {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
