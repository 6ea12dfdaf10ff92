"""Seeded synthetic weights and images for the peekvit hot path (test infrastructure).

The reference zero-initialises ``head.*``, ``class_tokens`` and every MHA bias
(reference ``models/vit.py:165,186-188``; ``nn.MultiheadAttention`` zero-inits
``in_proj_bias``/``out_proj.bias``), which makes random-init logits identically
zero and any parity check vacuous.  ``make_state_dict``
therefore draws *every* tensor of the checkpoint contract from a
seeded generator, in a fixed key order, so the same tensors can be loaded into the
reference modules (``load_state_dict(strict=True)`` in ``tests/golden/make_golden.py``)
and into ``peekvit_b200``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional

import torch

FAMILIES = ("vit", "rankvit", "residualvit", "adavit", "moevit", "eeresidualvit")


def _num_patches(cfg) -> int:
    return (cfg["image_size"] // cfg["patch_size"]) ** 2


def seq_length(family: str, cfg) -> int:
    """Length of ``encoder.pos_embedding`` (reference vit.py:162-171, residualvit.py:437-448)."""
    n = _num_patches(cfg) + cfg.get("num_class_tokens", 1)
    if family != "moevit":
        n += cfg.get("num_registers", 0)
    return n


def make_state_dict(family: str, cfg: Dict, seed: int = 4321, gate_std: float = 0.5) -> "OrderedDict[str, torch.Tensor]":
    """All parameters of one model family, fp32, drawn from ``torch.Generator(seed)``.

    Names and shapes follow the live ``state_dict()`` of the reference classes
    (SURVEY.md §8b): vit.py:160-188, rankvit.py:192-223, residualvit.py:437-493,
    adavit.py:300-324, moevit.py:240-265.
    """
    assert family in FAMILIES, family
    ee = family == "eeresidualvit"           # same blocks as ResidualViT + one early-exit head per layer (eeresidualvit.py:73-75)
    if ee:
        family = "residualvit"
    g = torch.Generator().manual_seed(seed)
    D, F, L = cfg["hidden_dim"], cfg["mlp_dim"], cfg["num_layers"]
    C, p = cfg["num_classes"], cfg["patch_size"]
    T = cfg.get("num_class_tokens", 1)
    R = cfg.get("num_registers", 0) if family != "moevit" else 0
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def normal(*shape, std=0.02, mean=0.0):
        return torch.randn(*shape, generator=g) * std + mean

    def uniform(*shape, bound):
        return (torch.rand(*shape, generator=g) * 2 - 1) * bound

    def linear(prefix, out_f, in_f, bias_std=0.02):
        sd[prefix + ".weight"] = uniform(out_f, in_f, bound=1.0 / math.sqrt(in_f))
        sd[prefix + ".bias"] = normal(out_f, std=bias_std)

    def layernorm(prefix):
        sd[prefix + ".weight"] = normal(D, std=0.02, mean=1.0)
        sd[prefix + ".bias"] = normal(D, std=0.02)

    def attention(prefix):
        sd[prefix + ".in_proj_weight"] = uniform(3 * D, D, bound=math.sqrt(6.0 / (4 * D)))
        sd[prefix + ".in_proj_bias"] = normal(3 * D, std=0.02)
        sd[prefix + ".out_proj.weight"] = uniform(D, D, bound=1.0 / math.sqrt(D))
        sd[prefix + ".out_proj.bias"] = normal(D, std=0.02)

    if family == "moevit":
        sd["class_token"] = normal(1, 1, D)
    else:
        sd["class_tokens"] = normal(1, T, D)
        if R > 0:
            sd["register_tokens"] = normal(1, R, D)
    if family == "residualvit":
        abt = cfg.get("add_budget_token", False)
        if abt in ("learnable", "learnable_interpolate"):
            sd["learnable_budget_token_1"] = normal(1, 1, D, std=1.0)
        if abt == "learnable_interpolate" or (ee and abt == "learnable"):      # eeresidualvit.py:212-216 always creates both
            sd["learnable_budget_token_2"] = normal(1, 1, D, std=1.0)
    fan_in = 3 * p * p
    sd["conv_proj.weight"] = normal(D, 3, p, p, std=math.sqrt(1.0 / fan_in)).clamp_(-2 * math.sqrt(1.0 / fan_in), 2 * math.sqrt(1.0 / fan_in))
    sd["conv_proj.bias"] = normal(D, std=0.02)
    sd["encoder.pos_embedding"] = normal(1, seq_length(family, cfg), D, std=0.02)

    for i in range(L):
        lp = f"encoder.layers.{i}"
        if family == "residualvit":
            skip = (cfg.get("residual_layers") or ["attention+mlp"] * L)[i]
            if skip in ("attention", "mlp", "attention+mlp"):
                # gate weights get a larger std so realised keep-fractions are not degenerate
                sd[lp + ".residual_gate.projection.weight"] = normal(1, D, std=gate_std / math.sqrt(D))
                sd[lp + ".residual_gate.projection.bias"] = normal(1, std=0.1)
        layernorm(lp + ".ln_1")
        if family == "moevit":
            ea = (cfg.get("attn_moes") or [1] * L)[i]
            linear(lp + ".self_attention.gating_network.gate", ea, D)
            if ea > 1:
                sd[lp + ".self_attention.gating_network.gate.weight"] *= 4.0
            for e in range(ea):
                attention(lp + f".self_attention.experts.{e}.self_attention")
        else:
            attention(lp + ".self_attention.self_attention")
        layernorm(lp + ".ln_2")
        if family == "moevit":
            em = (cfg.get("mlp_moes") or [1] * L)[i]
            linear(lp + ".mlp.gating_network.gate", em, D, bias_std=0.02)
            # spread router logits so the arg-max expert is not a near-tie everywhere
            sd[lp + ".mlp.gating_network.gate.weight"] *= 4.0
            for e in range(em):
                linear(lp + f".mlp.experts.{e}.fc1", F, D)
                linear(lp + f".mlp.experts.{e}.fc2", D, F)
        else:
            linear(lp + ".mlp.fc1", F, D)
            linear(lp + ".mlp.fc2", D, F)
        if family == "residualvit" and cfg.get("add_budget_token", False) == "learnable":
            sd[lp + ".budget_token_gate.weight"] = normal(1, D, std=1.0 / math.sqrt(D))
            sd[lp + ".budget_token_gate.bias"] = normal(1, std=0.1)
    layernorm("encoder.ln")
    if ee:
        for i in range(L):
            layernorm(f"encoder.early_exit_heads.{i}.0")
            linear(f"encoder.early_exit_heads.{i}.1", C, D)
    sd["head.weight"] = uniform(C, D, bound=1.0 / math.sqrt(D))
    sd["head.bias"] = normal(C, std=0.02)
    return sd


def synthetic_images(batch: int, image_size: int, seed: int = 1234, dtype=torch.float32) -> torch.Tensor:
    """``randn`` images: post-``T.Normalize`` Imagenette/ImageNet pixels are ~N(0,1)
    (reference data/imagenette.py:69-73)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, image_size, image_size, generator=g).to(dtype)


def calibrate_residual_gates(sd, cfg, target_keep: float, images: Optional[torch.Tensor] = None):
    """shift ``residual_gate.projection.bias`` per layer so the realised
    keep-fraction of a probe batch is ~``target_keep`` at budget ``target_keep``.  Returns a
    new state dict; the same tensors are then used by the oracle and the CUDA path."""
    from . import peekvit_oracle as po

    sd = OrderedDict((k, v.clone()) for k, v in sd.items())
    if images is None:
        images = synthetic_images(4, cfg["image_size"], seed=99)
    L = cfg["num_layers"]
    for i in range(L):
        key = f"encoder.layers.{i}.residual_gate.projection.bias"
        if key not in sd:
            continue
        lo, hi = -30.0, 30.0
        for _ in range(24):
            mid = 0.5 * (lo + hi)
            sd[key] = torch.tensor([mid])
            _, aux = po.residualvit_forward(sd, cfg, images, budget=target_keep, stop_after_layer=i)
            frac = float((aux["masks"][i] > 0).float().mean())
            if frac > target_keep:
                hi = mid
            else:
                lo = mid
        sd[key] = torch.tensor([0.5 * (lo + hi)])
    return sd
