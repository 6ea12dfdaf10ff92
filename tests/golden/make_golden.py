"""Generate the golden fixtures in this directory from the *imported reference*.

Run in the authoring container only (needs ``/root/reference``; it cannot travel to
the GPU box):

    python tests/golden/make_golden.py

For every model family it (1) builds the reference ``nn.Module`` from a small config,
(2) loads the oracle's seeded state dict with ``strict=True`` (this also pins the
checkpoint name/shape contract), (3) runs the reference forward on
seeded synthetic images, (4) checks the oracle restatement against it and
(5) stores the reference outputs (logits + published side-state) as ``<case>.npz``.
Weights and images are *not* stored: tests regenerate them from the recorded seeds
with ``oracle.weights`` (torch's CPU generator is deterministic for a fixed version).
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import weights as ow  # noqa: E402
from oracle import peekvit_oracle as po  # noqa: E402
from golden_cases import CASES, build_case, oracle_noise  # noqa: E402


def import_reference():
    tmp = tempfile.mkdtemp(prefix="peekvit_ref_")
    os.symlink("/root/reference", os.path.join(tmp, "peekvit"))
    sys.path.insert(0, tmp)
    # AViT hard-codes .cuda() (reference adavit.py:148-152,187); no-op it on CPU.
    torch.Tensor.cuda = lambda self, *a, **k: self
    from peekvit.models.vit import VisionTransformer
    from peekvit.models.rankvit import RankVisionTransformer
    from peekvit.models.residualvit import ResidualVisionTransformer
    from peekvit.models.adavit import AdaptiveVisionTransformer
    from peekvit.models.moevit import VisionTransformerMoE
    from peekvit.models.eeresidualvit import EEResidualVisionTransformer
    return {"eeresidualvit": EEResidualVisionTransformer, "vit": VisionTransformer, "rankvit": RankVisionTransformer,
            "residualvit": ResidualVisionTransformer, "adavit": AdaptiveVisionTransformer,
            "moevit": VisionTransformerMoE}


def stable_argsort_patch():
    """The reference's argsort is non-stable; fixtures are generated under the contract's
    tie rule (stable, lowest index first).  With continuous random scores there are no ties,
    so this does not change the reference's results — asserted below."""
    orig = torch.argsort
    seen = []

    def stable(x, dim=-1, descending=False, stable=False):
        r = orig(x, dim=dim, descending=descending, stable=True)
        seen.append(r)
        return r
    return orig, stable, seen


def main():
    classes = import_reference()
    torch.manual_seed(0)
    only = set(sys.argv[1:])                         # optional: regenerate just the named cases
    for name, case in CASES.items():
        if only and name not in only:
            continue
        fam, cfg, budget = case["family"], dict(case["cfg"]), case.get("budget")
        sd, images = build_case(case)
        model = classes[fam](**cfg)
        missing = model.load_state_dict(sd, strict=True)
        model.eval()
        if budget is not None:
            model.set_budget(budget)
        noise = None
        if case.get("noise") is not None:
            from peekvit.utils.utils import add_noise                    # the reference's own splice (utils/utils.py:162-191)
            add_noise(model, **case["noise"])
            noise = oracle_noise(case)                                   # same seed, drawn first: what the forward below draws
            torch.manual_seed(case["noise_seed"])
        out = {}
        with torch.no_grad():
            if fam == "rankvit":
                # confirm no ties -> default argsort == stable argsort
                orig, st, seen = stable_argsort_patch()
                logits_default = model(images)
                torch.argsort = st
                logits = model(images)
                torch.argsort = orig
                assert torch.equal(logits, logits_default)
            else:
                logits = model(images)
        if noise is not None:
            with torch.no_grad():
                if fam == "residualvit":
                    o_logits, aux = po.residualvit_forward(sd, cfg, images, budget, noise=noise)
                    clean, _ = po.residualvit_forward(sd, cfg, images, budget)
                else:
                    o_logits, aux = po.vit_forward(sd, cfg, images, noise=noise)
                    clean, _ = po.vit_forward(sd, cfg, images)
            assert (clean - logits).abs().max() > 1e-3, "the noise block had no effect: vacuous fixture"
        else:
            o_logits, aux = po.forward(fam, sd, cfg, images, budget)
        if fam == "eeresidualvit":
            # outputs are a list: one early exit per layer, then the final logits (eeresidualvit.py:355-357)
            assert len(logits) == cfg["num_layers"] + 1 == len(o_logits)
            for i, (r, o) in enumerate(zip(logits[:-1], o_logits[:-1])):
                e = (o - r).abs().max().item() / r.abs().max().item()
                assert e < 2e-5, (name, i, e)
                out[f"exit_{i}"] = r.numpy()
            logits, o_logits = logits[-1], o_logits[-1]
        out["logits"] = logits.numpy()
        err = (o_logits - logits).abs().max().item() / logits.abs().max().item()
        print(f"{name:28s} max|logit|={logits.abs().max():.3f} oracle-vs-reference rel err {err:.2e}")
        assert err < 2e-5, (name, err)
        if fam in ("residualvit", "eeresidualvit"):
            blocks = [b for b in model.encoder.layers if type(b).__name__ != "NoiseBlock"]     # fixture keys: block index
            for i, blk in enumerate(blocks):
                if getattr(blk, "mask", None) is not None:
                    out[f"mask_{i}"] = blk.mask.numpy()
                    assert torch.allclose(aux["masks"][i], blk.mask, atol=2e-5), (name, i)
                    print(f"    layer {i} keep-fraction {(blk.mask > 0).float().mean():.3f}")
        if fam == "adavit":
            out["rho_token"] = model.encoder.rho_token.numpy()
            out["counter_token"] = model.encoder.counter_token.numpy()
            assert torch.allclose(aux["rho_token"], model.encoder.rho_token, atol=1e-5)
            assert torch.equal(aux["counter_token"], model.encoder.counter_token)
            print("    counter_token mean", float(model.encoder.counter_token.mean()))
        if fam == "moevit":
            for i, blk in enumerate(model.encoder.layers):
                gp = getattr(blk.mlp, "gating_probs", None)
                if gp is not None:
                    out[f"mlp_gating_{i}"] = gp.argmax(-1).numpy().astype(np.int32)
                    assert torch.equal(aux["mlp_gating"][i], gp)
                gp = getattr(blk.self_attention, "gating_probs", None)
                if gp is not None:
                    out[f"attn_gating_{i}"] = gp.argmax(-1).numpy().astype(np.int32)
                    assert torch.equal(aux["attn_gating"][i], gp)
        if fam == "rankvit":
            for j, (i, idx) in enumerate(sorted(aux["kept"].items())):
                # the reference's own (stable) argsort output, first k entries
                assert torch.equal(seen[j][:, :idx.shape[1]], idx), (name, i)
                out[f"kept_{i}"] = seen[j][:, :idx.shape[1]].numpy().astype(np.int32)
            out["seq_lens"] = np.asarray(aux["seq_lens"], dtype=np.int32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    manifest_path = os.path.join(HERE, "MANIFEST.json")
    manifest = {}
    if os.path.exists(manifest_path):                 # keep what the other generators recorded (gradient fixtures)
        with open(manifest_path) as f:
            manifest = json.load(f)
    manifest.update({"torch": torch.__version__, "cases": sorted(CASES)})
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
