"""Golden gradients of the ResidualViT gate regime, from the *imported reference* in training mode.

Run in the authoring container only (needs ``/root/reference``):

    python tests/golden/make_finetune_residual.py

Builds the reference ``ResidualVisionTransformer`` of golden case ``residual_learnable_cal04`` (sigmoid gates, learnable budget
token, calibrated gate biases), puts it in ``train()`` mode with ``train_only_these_params(['gate', 'class', 'head',
'threshold', 'budget'])`` (train/train.py:99-100), fixes the per-image budgets the training forward would sample
(``_sample_budget``, residualvit.py:541-550), and runs ``loss = CrossEntropyLoss()(model(x), y) + 0.5 * mask regulariser;
loss.backward()``.  The regulariser restates ``solo_mse(per_layer=False, strict=True)`` (utils/losses.py:111-142, the
``MSELoss`` of configs/loss/crossentropy_mse.yaml in its strict form so that it is active on these masks; ``utils.losses``
itself needs hydra / omegaconf, absent here), weighted 0.5 so that its mask gradients matter next to the cross-entropy's.  Stored: logits, loss,
masks, and the gradient of every trainable parameter -> ``finetune_residual_learnable.npz``.  The oracle restatement is checked
against the same gradients through torch autograd.
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
from make_golden import import_reference  # noqa: E402
from oracle import peekvit_oracle as po  # noqa: E402

CASE = "residual_learnable_cal04"
BUDGETS = [0.3, 0.5, 0.7, 0.9]
WORDS = ["gate", "class", "head", "threshold", "budget"]
REG_WEIGHT = 0.5


def mask_regulariser(masks, budget):
    """solo_mse(per_layer=False, strict=True, skip_layers=[]) on a list of (B, N, 1) masks and (B,) budgets."""
    sp = torch.stack([m.mean(dim=(1, 2)) for m in masks]).mean()
    return ((sp - budget) ** 2).sum().mul(2 - budget).mean()


def main():
    refs = import_reference()
    case = CASES[CASE]
    sd, images = build_case(case)
    B = images.shape[0]
    labels = torch.arange(B) % case["cfg"]["num_classes"]
    budgets = torch.tensor(BUDGETS[:B])
    model = refs["residualvit"](**case["cfg"])
    model.load_state_dict(sd, strict=True)
    model.train()
    names = []
    for n, p in model.named_parameters():
        p.requires_grad = any(w in n for w in WORDS)
        if p.requires_grad:
            names.append(n)
    model._sample_budget = lambda n: budgets.clone()
    out = model(images)
    masks = [blk.mask for blk in model.encoder.layers]
    ce = torch.nn.functional.cross_entropy(out, labels)
    reg = mask_regulariser(masks, model.current_budget)
    loss = ce + REG_WEIGHT * reg
    loss.backward()
    store = {"logits": out.detach().numpy(), "loss": np.float32(loss.item()), "ce": np.float32(ce.item()),
             "budgets": budgets.numpy(), "labels": labels.numpy()}
    for i, m in enumerate(masks):
        store[f"mask.{i}"] = m.detach().numpy()
    grads = {}
    for n, p in model.named_parameters():
        if p.requires_grad and p.grad is not None:
            grads[n] = p.grad.detach().clone()
            store["grad." + n] = grads[n].numpy()
    print(f"reference: loss {loss.item():.6f} (ce {ce.item():.6f}, reg {reg.item():.6f}); {len(grads)} parameter gradients; "
          f"keep fractions {[round(float((m > 0).float().mean()), 3) for m in masks]}")
    # the oracle restatement through autograd must give the same numbers
    sdg = {k: v.clone() for k, v in sd.items()}
    for n in grads:
        sdg[n].requires_grad_(True)
    ologits, oaux = po.residualvit_forward(sdg, case["cfg"], images, budgets.view(B, 1, 1))
    oloss = torch.nn.functional.cross_entropy(ologits, labels) + REG_WEIGHT * mask_regulariser(
        [oaux["masks"][i] for i in sorted(oaux["masks"])], budgets)
    oloss.backward()
    worst = max(((sdg[n].grad - g).abs().max() / g.abs().max().clamp_min(1e-12)).item() for n, g in grads.items())
    print(f"oracle: loss {oloss.item():.6f}; worst gradient rel err vs reference {worst:.2e}; logits err "
          f"{(ologits - out).abs().max().item():.2e}")
    assert worst < 1e-4 and abs(oloss.item() - loss.item()) < 1e-5
    path = os.path.join(HERE, "finetune_residual_learnable.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
