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


def _check_model(model):
    if model.training:
        raise RuntimeError("peekvit_b200 implements the inference forward only; call model.eval() "
                           "(training / fine-tuning is a later row, SURVEY.md §8 f4)")
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("peekvit_b200 has no CPU path: move the model to a B200 with model.cuda()")
    return dev


def _check_images(model, x):
    torch._assert(x.dim() == 4 and x.shape[1] == 3, f"Expected (batch, 3, H, W) got {tuple(x.shape)}")
    torch._assert(x.shape[2] == model.image_size, f"Wrong image height! Expected {model.image_size} but got {x.shape[2]}!")
    torch._assert(x.shape[3] == model.image_size, f"Wrong image width! Expected {model.image_size} but got {x.shape[3]}!")


def _forward_chunk(model, fwd: engine.Forward, chunk: torch.Tensor, aux: Optional[dict]) -> torch.Tensor:
    family = model._family
    if family == "vit":
        return fwd.vit(chunk)
    if family == "rankvit":
        return fwd.rankvit(chunk, _rank_budgets(model), aux)
    raise NotImplementedError(f"{type(model).__name__}: forward for family {family!r} is not built yet")


def run(model, x: torch.Tensor, aux: Optional[dict] = None) -> torch.Tensor:
    """``model(images)`` with the images already on the model's device (validate/test.py:117-119)."""
    dev = _check_model(model)
    if x.device != dev:
        raise RuntimeError(f"input is on {x.device} but the model is on {dev}")
    _check_images(model, x)
    x = x.detach().to(torch.float32).contiguous()
    with torch.no_grad():
        fwd = engine.Forward(packed(model), workspace(model, dev))
        mb = int(getattr(model, "pk_micro_batch", DEFAULT_MICRO_BATCH))
        B = x.shape[0]
        out = torch.empty(B, model.num_classes, dtype=torch.float32, device=dev)
        for s in range(0, B, mb):
            chunk = x[s:s + mb]
            out[s:s + chunk.shape[0]].copy_(_forward_chunk(model, fwd, chunk, aux))
    return out


def run_host(model, x_host: torch.Tensor, out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
    """End-to-end entry for host-resident batches (the eval loop's ``batch.to(device); model(batch)``,
    validate/test.py:117-121, as one call): micro-batches are copied host->device on a side stream,
    double-buffered against compute, and the logits are returned in host memory."""
    dev = _check_model(model)
    if x_host.device.type != "cpu":
        raise RuntimeError("run_host expects a CPU (ideally pinned) tensor; use model(x) for device tensors")
    _check_images(model, x_host)
    if x_host.dtype != torch.float32 or not x_host.is_contiguous():
        x_host = x_host.to(torch.float32).contiguous()
    B, S = x_host.shape[0], model.image_size
    with torch.no_grad():
        ws = workspace(model, dev)
        fwd = engine.Forward(packed(model), ws)
        mb = min(int(getattr(model, "pk_micro_batch", DEFAULT_MICRO_BATCH)), max(B, 1))
        st = _state(model)
        if "copy_stream" not in st:
            st["copy_stream"] = torch.cuda.Stream(device=dev)
            st["ready"] = [torch.cuda.Event(), torch.cuda.Event()]
            st["free"] = [torch.cuda.Event(), torch.cuda.Event()]
        copy_stream, ready, free = st["copy_stream"], st["ready"], st["free"]
        bufs = [ws.get("img_stage0", (mb, 3, S, S), torch.float32), ws.get("img_stage1", (mb, 3, S, S), torch.float32)]
        out = ws.get("logits_all", (B, model.num_classes), torch.float32)
        cur = torch.cuda.current_stream(dev)
        for i in range(2):
            free[i].record(cur)
        for i, s in enumerate(range(0, B, mb)):
            n = min(mb, B - s)
            slot = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[slot])
                bufs[slot][:n].copy_(x_host[s:s + n], non_blocking=True)
                ready[slot].record(copy_stream)
            cur.wait_event(ready[slot])
            out[s:s + n].copy_(_forward_chunk(model, fwd, bufs[slot][:n], None))
            free[slot].record(cur)
        if out_host is None:
            out_host = torch.empty(B, model.num_classes, dtype=torch.float32, pin_memory=True)
        out_host.copy_(out, non_blocking=True)
        cur.synchronize()
    return out_host
