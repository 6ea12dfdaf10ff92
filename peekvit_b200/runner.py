"""``model(images)`` for the drop-in modules: validation, weight prepack cache, micro-batching.

Mirrors what the reference's callers expect from ``forward`` (validate/test.py:117-121):
``(B,3,S,S)`` float images on the model's device in, ``(B,num_classes)`` fp32 logits out, Python
exceptions on misuse.  No autograd, no CPU path.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import engine

DEFAULT_MICRO_BATCH = int(os.environ.get("PEEKVIT_B200_MICRO_BATCH", "128"))


def _state(model):
    st = model.__dict__.get("_pk_state")
    if st is None:
        st = {"fp": None, "pm": None, "ws": None}
        model.__dict__["_pk_state"] = st
    return st


def packed(model) -> engine.PackedModel:
    """bf16/fp32 prepacked weights, rebuilt whenever a parameter changed (load_state_dict,
    optimiser step, .to(), del encoder.layers[i] …)."""
    st = _state(model)
    fp = engine.params_fingerprint(model)
    if st["fp"] != fp:
        st["pm"] = engine.pack_model(model, model._family)
        st["fp"] = fp
    return st["pm"]


def workspace(model, device) -> engine.Workspace:
    st = _state(model)
    if st["ws"] is None or st["ws"].device != device:
        st["ws"] = engine.Workspace(device)
    return st["ws"]


def _rank_budgets(model) -> Dict[int, float]:
    return {i: float(blk.current_budget) for i, blk in enumerate(model.encoder.layers) if hasattr(blk, "current_budget")}


def run(model, x: torch.Tensor, aux: Optional[dict] = None) -> torch.Tensor:
    if model.training:
        raise RuntimeError("peekvit_b200 implements the inference forward only; call model.eval() "
                           "(training / fine-tuning is a later row, SURVEY.md §8 f4)")
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("peekvit_b200 has no CPU path: move the model to a B200 with model.cuda()")
    if x.device != dev:
        raise RuntimeError(f"input is on {x.device} but the model is on {dev}")
    torch._assert(x.dim() == 4 and x.shape[1] == 3, f"Expected (batch, 3, H, W) got {tuple(x.shape)}")
    torch._assert(x.shape[2] == model.image_size, f"Wrong image height! Expected {model.image_size} but got {x.shape[2]}!")
    torch._assert(x.shape[3] == model.image_size, f"Wrong image width! Expected {model.image_size} but got {x.shape[3]}!")
    x = x.detach().to(torch.float32).contiguous()
    with torch.no_grad():
        pm = packed(model)
        ws = workspace(model, dev)
        fwd = engine.Forward(pm, ws)
        mb = int(getattr(model, "pk_micro_batch", DEFAULT_MICRO_BATCH))
        B = x.shape[0]
        out = torch.empty(B, model.num_classes, dtype=torch.float32, device=dev)
        family = model._family
        for s in range(0, B, mb):
            chunk = x[s:s + mb]
            if family == "vit":
                logits = fwd.vit(chunk)
            elif family == "rankvit":
                logits = fwd.rankvit(chunk, _rank_budgets(model), aux)
            else:
                raise NotImplementedError(f"{type(model).__name__}: forward for family {family!r} is not built yet")
            out[s:s + chunk.shape[0]].copy_(logits)
    return out
