"""Golden gradients of the class-token + head regime, from the *imported reference* in training mode.

Run in the authoring container only (needs ``/root/reference``):

    python tests/golden/make_finetune_vit.py

Builds the reference ``VisionTransformer`` of golden case ``vit_d128_regs`` (two class tokens, registers) and the reference
``RankVisionTransformer`` of case ``rankvit_b05``, puts them in ``train()`` mode with ``train_only_these_params(['gate', 'class',
'head', 'threshold', 'budget'])`` (train/train.py:99-100, models/topology.py:128-158) and runs ``loss =
CrossEntropyLoss()(model(x), y); loss.backward()`` (train/train.py:105-113).  Stored per case: logits, loss and the gradients
of ``class_tokens`` / ``head.weight`` / ``head.bias`` -> ``finetune_<case>.npz``.  The oracle restatement is checked against
the same gradients through torch autograd.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from golden_cases import CASES, build_case  # noqa: E402
from make_golden import import_reference, stable_argsort_patch  # noqa: E402
from oracle import peekvit_oracle as po  # noqa: E402

WORDS = ["gate", "class", "head", "threshold", "budget"]


def main():
    refs = import_reference()
    stable_argsort_patch()
    for name in ("vit_d128_regs", "rankvit_b05"):
        case = CASES[name]
        sd, images = build_case(case)
        B = images.shape[0]
        labels = (torch.arange(B) * 3 + 1) % case["cfg"]["num_classes"]
        model = refs[case["family"]](**case["cfg"])
        model.load_state_dict(sd, strict=True)
        model.train()
        for n, p in model.named_parameters():
            p.requires_grad = any(w in n for w in WORDS)
        if case.get("budget") is not None:
            model.set_budget(case["budget"])
        out = model(images)
        loss = torch.nn.functional.cross_entropy(out, labels)
        loss.backward()
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.requires_grad and p.grad is not None}
        assert sorted(grads) == ["class_tokens", "head.bias", "head.weight"], sorted(grads)
        store = {"logits": out.detach().numpy(), "loss": np.float32(loss.item()), "labels": labels.numpy()}
        for n, g in grads.items():
            store["grad." + n] = g.numpy()
        sdg = {k: v.clone() for k, v in sd.items()}
        for n in grads:
            sdg[n].requires_grad_(True)
        if case["family"] == "vit":
            ologits, _ = po.vit_forward(sdg, case["cfg"], images)
        else:
            ologits, _ = po.rankvit_forward(sdg, case["cfg"], images, case["budget"])
        oloss = torch.nn.functional.cross_entropy(ologits, labels)
        oloss.backward()
        worst = max(((sdg[n].grad.view_as(g) - g).abs().max() / g.abs().max()).item() for n, g in grads.items())
        print(f"{name}: reference loss {loss.item():.6f}, oracle {oloss.item():.6f}, worst gradient rel err of the oracle {worst:.2e}")
        assert worst < 1e-4 and abs(oloss.item() - loss.item()) < 1e-5
        path = os.path.join(HERE, f"finetune_{name}.npz")
        np.savez_compressed(path, **store)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
