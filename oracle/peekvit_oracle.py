"""Functional CPU restatement of the peekvit encoder forward (test infrastructure).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Every function works on a plain
``state_dict`` mapping (``oracle.weights.make_state_dict``) plus a config dict and
restates the *reference semantics* (dense, masked — the reference never skips
work) with ``torch`` CPU ops in fp32 (or fp64 when the inputs are
fp64).  The arithmetic itself lives in PyTorch in the reference too
(``nn.Conv2d``, ``nn.LayerNorm``, ``nn.MultiheadAttention``, ``nn.Linear``,
``F.gelu``, ``torch.norm`` …, pinned ``torch>=2.1.2`` in reference
``requirements.txt:1-6``; torch 2.11.0 here), so the restatement is anchored on the
reference's call sites, cited per function, and pinned by
``tests/golden/*.npz`` generated from the imported reference.

The one place the oracle is *stricter* than the reference: token ranking uses a
stable descending sort (ties -> lowest index), the contract fixed by
BASELINE.json's north_star; the reference's ``torch.argsort`` is unspecified on
ties.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- primitives
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    """``nn.LayerNorm(hidden_dim)``: reference vit.py:37,42,88 (eps 1e-5) and
    residualvit.py:117,122 (eps 1e-6)."""
    return F.layer_norm(x, (x.shape[-1],), w.to(x.dtype), b.to(x.dtype), eps)


def mha(x: Tensor, sd, prefix: str, num_heads: int) -> Tensor:
    """``SelfAttention.forward`` (reference blocks.py:88-95) -> ``nn.MultiheadAttention``
    (batch_first, packed in-proj rows ``[q;k;v]``, heads are contiguous ``dh`` chunks,
    ``q`` pre-scaled by ``1/sqrt(dh)``, softmax over keys, out-proj).  The head-averaged
    attention weights the reference also materialises (``need_weights=True``,
    blocks.py:94) are discarded there and not produced here."""
    B, N, D = x.shape
    dh = D // num_heads
    dt = x.dtype
    qkv = F.linear(x, sd[prefix + ".in_proj_weight"].to(dt), sd[prefix + ".in_proj_bias"].to(dt))
    q, k, v = qkv.split(D, dim=-1)
    q = q.reshape(B, N, num_heads, dh).transpose(1, 2) * (dh ** -0.5)
    k = k.reshape(B, N, num_heads, dh).transpose(1, 2)
    v = v.reshape(B, N, num_heads, dh).transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, N, D)
    return F.linear(o, sd[prefix + ".out_proj.weight"].to(dt), sd[prefix + ".out_proj.bias"].to(dt))


def mlp(x: Tensor, sd, prefix: str) -> Tensor:
    """``MLP.forward``: fc1 -> exact (erf) GELU -> fc2 (reference blocks.py:74-84)."""
    dt = x.dtype
    h = F.gelu(F.linear(x, sd[prefix + ".fc1.weight"].to(dt), sd[prefix + ".fc1.bias"].to(dt)))
    return F.linear(h, sd[prefix + ".fc2.weight"].to(dt), sd[prefix + ".fc2.bias"].to(dt))


def patch_embed(images: Tensor, sd, cfg) -> Tensor:
    """``_process_input``: Conv2d(3->D, k=s=p)+bias, flatten, transpose
    (reference vit.py:203-222; identical copies residualvit.py:508-527,
    rankvit.py:235-254, adavit.py:336-355, moevit.py:274-293)."""
    n, c, h, w = images.shape
    p = cfg["patch_size"]
    assert h == cfg["image_size"] and w == cfg["image_size"], "Wrong image size"
    dt = images.dtype
    x = F.conv2d(images, sd["conv_proj.weight"].to(dt), sd["conv_proj.bias"].to(dt), stride=p)
    return x.reshape(n, cfg["hidden_dim"], (h // p) * (w // p)).permute(0, 2, 1)


def vit_block(x: Tensor, sd, lp: str, num_heads: int, eps: float = 1e-5,
              attn_prefix: Optional[str] = None, mlp_prefix: Optional[str] = None) -> Tensor:
    """``ViTBlock.forward``: ``x = attn(ln_1(in)) + in; out = x + mlp(ln_2(x))``
    (reference vit.py:45-55; dropout is identity in eval)."""
    a = mha(layer_norm(x, sd[lp + ".ln_1.weight"], sd[lp + ".ln_1.bias"], eps), sd,
            attn_prefix or (lp + ".self_attention.self_attention"), num_heads)
    x = a + x
    y = mlp(layer_norm(x, sd[lp + ".ln_2.weight"], sd[lp + ".ln_2.bias"], eps), sd, mlp_prefix or (lp + ".mlp"))
    return x + y


def _tokens(images: Tensor, sd, cfg, cls_key: str = "class_tokens") -> Tensor:
    """Patch tokens with registers then class tokens prepended: ``[cls, regs, patches]``
    (reference vit.py:226-236)."""
    x = patch_embed(images, sd, cfg)
    n = x.shape[0]
    if cfg.get("num_registers", 0) > 0 and "register_tokens" in sd:
        x = torch.cat([sd["register_tokens"].to(x.dtype).expand(n, -1, -1), x], dim=1)
    return torch.cat([sd[cls_key].to(x.dtype).expand(n, -1, -1), x], dim=1)


def _head(x: Tensor, sd, cfg) -> Tensor:
    """Final LN (all tokens in the reference, only cls rows consumed), class-token **sum**,
    linear head (reference vit.py:95,242-246)."""
    dt = x.dtype
    x = layer_norm(x, sd["encoder.ln.weight"], sd["encoder.ln.bias"], 1e-5)
    x = x[:, 0:cfg.get("num_class_tokens", 1)].sum(dim=1)
    return F.linear(x, sd["head.weight"].to(dt), sd["head.bias"].to(dt))


def stable_topk_desc(scores: Tensor, k: int) -> Tensor:
    """Indices of the ``k`` largest scores per row in descending order, ties -> lowest
    index (north_star contract; reference rankvit.py:67 uses a non-stable argsort)."""
    return torch.argsort(scores, dim=-1, descending=True, stable=True)[..., :k]


def noise_block(x: Tensor, noise: Dict) -> Tensor:
    """``NoiseBlock.forward`` (reference blocks.py:159-170) with the random draw passed in: ``noise['snr_db']`` +
    ``noise['noise']`` (the ``randn_like`` tensor) -> ``forward_snr`` (:117-131; 0 dB means no noise there), or
    ``noise['prob']`` + ``noise['perm']`` (the ``randperm`` result) -> ``forward_token_drop`` (:141-157)."""
    if noise.get("snr_db") is not None:
        if noise["snr_db"] == 0:
            return x
        power = torch.mean(x ** 2, dim=-1, keepdim=True)
        return x + noise["noise"].to(x.dtype) * torch.sqrt(power / (10 ** (noise["snr_db"] / 10)))
    if noise["prob"] == 0:
        return x
    keep = torch.ones_like(x)
    keep[:, noise["perm"][:int(noise["prob"] * x.shape[1])], :] = 0
    return x * keep


# --------------------------------------------------------------------------- plain ViT
def vit_forward(sd, cfg, images: Tensor, noise: Optional[Dict] = None) -> Tuple[Tensor, Dict]:
    """``VisionTransformer.forward`` (reference vit.py:224-248) with ``ViTEncoder.forward``
    (vit.py:90-95): ``+pos_embedding``, L blocks, LN, cls sum, head.  ``noise`` (optional) describes a NoiseBlock spliced in
    before block ``noise['layer']`` by ``add_noise`` (reference utils/utils.py:162-191)."""
    x = _tokens(images, sd, cfg)
    x = x + sd["encoder.pos_embedding"].to(x.dtype)
    for i in range(cfg["num_layers"]):
        if noise is not None and noise["layer"] == i:
            x = noise_block(x, noise)
        x = vit_block(x, sd, f"encoder.layers.{i}", cfg["num_heads"])
    if noise is not None and noise["layer"] >= cfg["num_layers"]:
        x = noise_block(x, noise)
    return _head(x, sd, cfg), {}


# --------------------------------------------------------------------------- RankViT
def rank_budget_for_layer(budget: Union[float, Sequence[float]], layer: int) -> float:
    """``RankVisionTransformer.set_budget``: scalar, or list indexed by *layer index*
    (reference rankvit.py:283-288)."""
    return float(budget[layer]) if isinstance(budget, (list, tuple)) else float(budget)


def rank_keep_count(n_tokens: int, budget: float) -> int:
    """``math.ceil(n * current_budget)`` on the *current* patch-token count, so ratios
    compound across rank layers (reference rankvit.py:74)."""
    return math.ceil(n_tokens * budget)


def rankvit_forward(sd, cfg, images: Tensor, budget: Union[float, Sequence[float]] = 1.0,
                    forced_kept: Optional[Dict[int, Tensor]] = None) -> Tuple[Tensor, Dict]:
    """``RankVisionTransformer.forward`` (reference rankvit.py:256-280); rank layers run
    ``sort_and_drop`` (rankvit.py:55-77) on the block's pre-LN input when
    ``current_budget != 1`` (rankvit.py:85-88): L2 norm over D, descending sort, keep the
    first ``ceil(n*b)``, kept tokens in descending-norm order after the class token.
    ``forced_kept`` (test hook) substitutes given index tensors for the selection of the listed
    layers, to compare logits *given identical selections*: top-k is discontinuous, so a bf16-level
    score perturbation can legitimately swap tokens at the cut."""
    rank_layers = list(cfg["rankvit_layers"])
    x = _tokens(images, sd, cfg)
    x = x + sd["encoder.pos_embedding"].to(x.dtype)
    aux = {"scores": {}, "kept": {}, "seq_lens": []}
    for i in range(cfg["num_layers"]):
        if i in rank_layers:
            b = rank_budget_for_layer(budget, i)
            if b != 1:
                cls, tok = x[:, 0:1], x[:, 1:]
                scores = torch.norm(tok, dim=-1)                       # rankvit.py:63
                k = rank_keep_count(tok.shape[1], b)                   # rankvit.py:74
                idx = stable_topk_desc(scores, k)                      # rankvit.py:67 (+tie rule)
                if forced_kept is not None and i in forced_kept:
                    assert forced_kept[i].shape == idx.shape
                    idx = forced_kept[i].to(torch.int64)
                kept = torch.gather(tok, 1, idx.unsqueeze(-1).expand(-1, -1, tok.shape[-1]))  # :71,75
                x = torch.cat([cls, kept], dim=1)                      # :77
                aux["scores"][i] = scores
                aux["kept"][i] = idx
        aux["seq_lens"].append(x.shape[1])
        x = vit_block(x, sd, f"encoder.layers.{i}", cfg["num_heads"])
    return _head(x, sd, cfg), aux


# --------------------------------------------------------------------------- ResidualViT
def residual_gate(img: Tensor, sd, lp: str, cfg, budget=None, threshold=None) -> Tensor:
    """``ResidualGate.forward`` (reference residualvit.py:47-74): Linear(D,1) ->
    ``SigmoidWithTemp`` ``sigmoid(x/temp + bias)`` (blocks.py:62-69) -> ``relu(mask-(1-budget))``
    (:62) or ``relu(mask-threshold)`` (:65) or the fixed ``gate_threshold`` (:69).
    ``gate_type='gumbel'`` in eval is ``round(sigmoid(x))`` (blocks.py:55-57)."""
    dt = img.dtype
    logit = F.linear(img, sd[lp + ".residual_gate.projection.weight"].to(dt),
                     sd[lp + ".residual_gate.projection.bias"].to(dt))
    if cfg.get("gate_type", "gumbel") == "sigmoid":
        m = torch.sigmoid(logit / cfg.get("gate_temp", 1.0) + cfg.get("gate_bias", 10.0))
        if budget is not None:
            m = F.relu(m - (1 - budget))
        elif threshold is not None:
            m = F.relu(m - threshold)
        else:
            m = F.relu(m - cfg.get("gate_threshold", 0.5))
        return m
    assert budget is None, "Gumbel gate does not support budget"
    return torch.round(torch.sigmoid(logit))


def _residual_plain(x: Tensor, sd, lp: str, H: int, mask: Union[Tensor, float]) -> Tensor:
    """``ResidualViTBlock.plain_forward`` (reference residualvit.py:249-260), LN eps 1e-6."""
    a = mask * layer_norm(x, sd[lp + ".ln_1.weight"], sd[lp + ".ln_1.bias"], 1e-6)
    a = mask * mha(a, sd, lp + ".self_attention.self_attention", H)
    x1 = a + x
    y = mask * layer_norm(x1, sd[lp + ".ln_2.weight"], sd[lp + ".ln_2.bias"], 1e-6)
    return x1 + mlp(y, sd, lp + ".mlp")


def residualvit_forward(sd, cfg, images: Tensor, budget: float, stop_after_layer: Optional[int] = None,
                        early_exits: Optional[List[Tensor]] = None, noise: Optional[Dict] = None,
                        forced_masks: Optional[Dict[int, Tensor]] = None) -> Tuple[Tensor, Dict]:
    """``ResidualVisionTransformer.forward`` in eval (reference residualvit.py:587-616):
    budget token built by ``_add_budget_token`` (:552-585) and appended last; encoder adds
    ``pos_embedding`` to all but the budget token (:335-348); blocks dispatch on ``skip``
    (:263-273): ``'attention+mlp'`` (:197-244), ``'attention'`` (:130-157), ``'mlp'`` (:160-194), ``None``/``'none'``
    (plain).  Tensor shapes follow the reference literally, so the combinations it cannot run (``'attention'`` with a
    budget token, ``add_input`` with a budget token, ``'attention+mlp'`` without one) fail here with the same shape
    errors.  ``noise`` describes a NoiseBlock spliced in front of block ``noise['layer']`` (utils/utils.py:162-191).
    ``forced_masks`` (test aid, like ``moevit_forward``'s forced routing): gate values to use instead of the computed ones, so a
    path whose hard 0/1 gumbel decisions flipped at a near-tie can be checked given identical decisions."""
    abt = cfg.get("add_budget_token", False)
    L, H, D = cfg["num_layers"], cfg["num_heads"], cfg["hidden_dim"]
    skips = cfg.get("residual_layers") or ["attention+mlp"] * L
    add_input = cfg.get("add_input", False)
    x = _tokens(images, sd, cfg)
    n, dt = x.shape[0], x.dtype
    x = x + sd["encoder.pos_embedding"].to(dt)
    if abt:
        assert budget is not None, "Budget token not set. Call set_budget() before forward()"
        cur = torch.as_tensor(budget, dtype=dt)
        if abt == "learnable":
            bt = sd["learnable_budget_token_1"].to(dt).expand(n, -1, -1) * cur          # :572-576
        elif abt == "learnable_interpolate":
            bt = (sd["learnable_budget_token_1"].to(dt) * cur + sd["learnable_budget_token_2"].to(dt) * (1 - cur)).expand(n, -1, -1)
        else:
            bt = torch.full((n, 1, D), float(budget), dtype=dt)                          # :581-583
        x = torch.cat([x, bt], dim=1)
    aux = {"masks": {}, "thresholds": {}}
    ln = lambda t, lp, k: layer_norm(t, sd[f"{lp}.{k}.weight"], sd[f"{lp}.{k}.bias"], 1e-6)
    for i in range(L):
        if noise is not None and noise["layer"] == i:
            x = noise_block(x, noise)
        lp = f"encoder.layers.{i}"
        skip = skips[i]
        # the model builds its encoder without num_class_tokens / num_registers (:452-467): blocks see ONE special token
        if skip == "attention+mlp":
            special, img = x[:, :1], x[:, 1:]
            cur_b = thr = None
            if abt:
                btok, img = img[:, -1:], img[:, :-1]
                cur_b = btok.mean()                       # whole-batch mean (:208)
            if abt == "learnable":
                thr = torch.sigmoid(F.linear(btok, sd[lp + ".budget_token_gate.weight"].to(dt),
                                             sd[lp + ".budget_token_gate.bias"].to(dt)))  # :212
                cur_b = None
                aux["thresholds"][i] = thr
            mask = residual_gate(img, sd, lp, cfg, budget=cur_b, threshold=thr)          # :217
            if forced_masks is not None:
                mask = forced_masks[i].to(dt)
            aux["masks"][i] = mask
            parts = [special, mask * img]
            if abt:
                parts.append(btok)
            fm = torch.cat([torch.ones(n, 1, 1, dtype=dt), mask, torch.ones(n, 1, 1, dtype=dt)], dim=1)   # :230-235
            y = _residual_plain(torch.cat(parts, dim=1), sd, lp, H, fm)
            if add_input:
                y = y + torch.cat([torch.zeros_like(special), img * (1 - mask)], dim=1)  # :239-242
            x = y
        elif skip == "attention":
            special, img = x[:, :1], x[:, 1:]
            cur_b = None
            if abt:
                btok, img = img[:, -1:], img[:, :-1]
                cur_b = btok.mean()                                                       # :142
            mask = residual_gate(img, sd, lp, cfg, budget=cur_b)
            if forced_masks is not None:
                mask = forced_masks[i].to(dt)
            aux["masks"][i] = mask
            mi = torch.cat([special, mask * img], dim=1)                                  # :145-148 (the budget token is not put back)
            x1 = mha(ln(mi, lp, "ln_1"), sd, lp + ".self_attention.self_attention", H) + x   # :150-153
            x = mlp(ln(x1, lp, "ln_2"), sd, lp + ".mlp")                                  # :155-157: no residual around the MLP
        elif skip == "mlp":
            x1 = mha(ln(x, lp, "ln_1"), sd, lp + ".self_attention.self_attention", H) + x    # :162-165
            special, img = x1[:, :1], x1[:, 1:]
            cur_b = None
            if abt:
                btok, img = img[:, -1:], img[:, :-1]
                cur_b = btok.mean()                                                       # :177
            mask = residual_gate(img, sd, lp, cfg, budget=cur_b)
            if forced_masks is not None:
                mask = forced_masks[i].to(dt)
            aux["masks"][i] = mask
            parts = [special, mask * img] + ([btok] if abt else [])
            y = mlp(ln(torch.cat(parts, dim=1), lp, "ln_2"), sd, lp + ".mlp")             # :186-187: no residual
            if add_input:
                y = y + torch.cat([torch.zeros_like(special), img * (1 - mask)], dim=1)   # :189-192
            x = y
        elif skip in (None, "none"):
            x = _residual_plain(x, sd, lp, H, 1.0)
        else:
            raise ValueError(f"unknown skip mode {skip!r}")
        if stop_after_layer is not None and i == stop_after_layer:
            return x, aux
        if early_exits is not None:
            hp = f"encoder.early_exit_heads.{i}"                           # eeresidualvit.py:73-75,94
            e = layer_norm(x[:, 0:1], sd[hp + ".0.weight"], sd[hp + ".0.bias"], 1e-5)
            early_exits.append(F.linear(e, sd[hp + ".1.weight"].to(dt), sd[hp + ".1.bias"].to(dt)).squeeze())
    return _head(x, sd, cfg), aux


def eeresidualvit_forward(sd, cfg, images: Tensor, budget: float) -> Tuple[List[Tensor], Dict]:
    """``EEResidualVisionTransformer.forward`` in eval (reference eeresidualvit.py:334-358): the ResidualViT forward
    (same blocks, budget token appended last, :278-331) plus, after every layer, ``early_exit_heads[i]`` =
    LayerNorm(1e-5) -> Linear applied to the first class token and squeezed (:91-96; the encoder is built with its default
    ``num_class_tokens=1``, :193-209); returns ``early_exits + [final logits]``.  A falsy budget (None or 0) raises
    ``ValueError`` (:308-309)."""
    if not budget:
        raise ValueError("Budget token not set. Call set_budget() before forward() to evaluate the model on a chosen budget.")
    exits: List[Tensor] = []
    logits, aux = residualvit_forward(sd, cfg, images, budget, early_exits=exits)
    return exits + [logits], aux


# --------------------------------------------------------------------------- AdaViT (A-ViT)
def avit_forward(sd, cfg, images: Tensor) -> Tuple[Tensor, Dict]:
    """``AdaptiveVisionTransformer.forward`` (reference adavit.py:357-381) with the ACT loop
    ``AViTEncoder.forward_features_act_token`` (adavit.py:140-219) and the masked block
    ``AViTBlock.forward_act`` (adavit.py:53-80)."""
    L, H = cfg["num_layers"], cfg["num_heads"]
    eps = cfg.get("eps", 0.01)
    scale, center = cfg.get("gate_scale", 10), cfg.get("gate_center", 30)
    x = _tokens(images, sd, cfg)
    dt = x.dtype
    x = x + sd["encoder.pos_embedding"].to(dt)
    bs, N, _ = x.shape
    c = torch.zeros(bs, N, dtype=dt)
    R = torch.ones(bs, N, dtype=dt)
    mask = torch.ones(bs, N, dtype=dt)
    rho = torch.zeros(bs, N, dtype=dt)
    counter = torch.ones(bs, N, dtype=dt)
    output = None
    out = x
    aux = {"halting_score_layer": [], "active": []}
    for i in range(L):
        lp = f"encoder.layers.{i}"
        aux["active"].append(mask.clone())
        out = out * mask.view(bs, N, 1)                                   # :170
        m3 = mask.view(bs, N, 1)
        a = mha(layer_norm(out * m3, sd[lp + ".ln_1.weight"], sd[lp + ".ln_1.bias"], 1e-5) * m3,
                sd, lp + ".self_attention.self_attention", H)            # :69
        xb = out + a
        xb = xb + mlp(layer_norm(xb * m3, sd[lp + ".ln_2.weight"], sd[lp + ".ln_2.bias"], 1e-5) * m3, sd, lp + ".mlp")  # :70
        h = torch.sigmoid(xb[:, :, 0] * scale - center)                    # :74
        aux["halting_score_layer"].append(h[1:].mean())                    # :176 (slices the batch dim)
        out = xb.clone()
        block_output = xb * m3                                             # :183
        if i == L - 1:
            h = torch.ones(bs, N, dtype=dt)                                # :186-187
        c = c + h                                                          # :190
        rho = rho + mask
        reached = (c > 1 - eps).to(dt) * mask                              # :195-196
        delta1 = block_output * R.view(bs, N, 1) * reached.view(bs, N, 1)
        rho = rho + R * reached
        not_reached = (c < 1 - eps).to(dt)                                 # :202-203
        R = R - not_reached * h
        delta2 = block_output * h.view(bs, N, 1) * not_reached.view(bs, N, 1)
        counter = counter + not_reached
        mask = (c < 1 - eps).to(dt)                                        # :210
        output = delta1 + delta2 if output is None else output + (delta1 + delta2)
    aux.update(rho_token=rho, counter_token=counter)
    return _head(output, sd, cfg), aux


# --------------------------------------------------------------------------- MoE ViT
def _moe_route(x: Tensor, sd, prefix: str) -> Tensor:
    """``TopKGate`` in eval: ``one_hot(argmax(Linear(D,E)(x)))`` (reference moevit.py:23-32,
    blocks.py:23-25)."""
    dt = x.dtype
    scores = F.linear(x, sd[prefix + ".gate.weight"].to(dt), sd[prefix + ".gate.bias"].to(dt))
    return F.one_hot(scores.argmax(dim=-1), num_classes=scores.shape[-1]).to(dt)


def moevit_forward(sd, cfg, images: Tensor, forced_mlp_expert: Optional[Dict[int, Tensor]] = None) -> Tuple[Tensor, Dict]:
    """``VisionTransformerMoE.forward`` (reference moevit.py:295-312; readout ``x[:, 0]``) with
    ``ViTBlockMoE`` (:131-141), ``MLPMoE.forward_moe`` (:49-61) and ``AttentionMoE`` (:71-102):
    every expert is evaluated densely and combined with the one-hot gate; a single expert
    bypasses the gate (:45-47,:64-67).  ``forced_mlp_expert`` (test hook, like ``forced_kept`` of the RankViT forward)
    substitutes given expert indices (B, N) for the arg-max routing of the listed layers, to compare logits *given
    identical routing*: arg-max is discontinuous, a bf16-level perturbation of a near-tied router score moves a token
    to another expert."""
    L, H = cfg["num_layers"], cfg["num_heads"]
    mlp_moes = cfg.get("mlp_moes") or [1] * L
    attn_moes = cfg.get("attn_moes") or [1] * L
    x = patch_embed(images, sd, cfg)
    dt = x.dtype
    x = torch.cat([sd["class_token"].to(dt).expand(x.shape[0], -1, -1), x], dim=1)
    x = x + sd["encoder.pos_embedding"].to(dt)
    aux = {"mlp_gating": {}, "attn_gating": {}}
    for i in range(L):
        lp = f"encoder.layers.{i}"
        a_in = layer_norm(x, sd[lp + ".ln_1.weight"], sd[lp + ".ln_1.bias"], 1e-5)
        if attn_moes[i] == 1:
            a = mha(a_in, sd, lp + ".self_attention.experts.0.self_attention", H)
        else:
            gp = _moe_route(a_in, sd, lp + ".self_attention.gating_network")
            aux["attn_gating"][i] = gp
            outs = torch.stack([mha(a_in, sd, lp + f".self_attention.experts.{e}.self_attention", H)
                                for e in range(attn_moes[i])], dim=0)
            a = torch.einsum("ebsd,bse->bsd", outs, gp)
        x = a + x
        m_in = layer_norm(x, sd[lp + ".ln_2.weight"], sd[lp + ".ln_2.bias"], 1e-5)
        if mlp_moes[i] == 1:
            y = mlp(m_in, sd, lp + ".mlp.experts.0")
        else:
            gp = _moe_route(m_in, sd, lp + ".mlp.gating_network")
            if forced_mlp_expert is not None and i in forced_mlp_expert:
                gp = F.one_hot(forced_mlp_expert[i].to(torch.int64), num_classes=mlp_moes[i]).to(dt)
            aux["mlp_gating"][i] = gp
            outs = torch.stack([mlp(m_in, sd, lp + f".mlp.experts.{e}") for e in range(mlp_moes[i])], dim=0)
            y = torch.einsum("ebsd,bse->bsd", outs, gp)
        x = x + y
    x = layer_norm(x, sd["encoder.ln.weight"], sd["encoder.ln.bias"], 1e-5)[:, 0]
    return F.linear(x, sd["head.weight"].to(dt), sd["head.bias"].to(dt)), aux


# --------------------------------------------------------------------------- dispatch / accounting
def forward(family: str, sd, cfg, images: Tensor, budget=None) -> Tuple[Tensor, Dict]:
    with torch.no_grad():
        if family == "vit":
            return vit_forward(sd, cfg, images)
        if family == "rankvit":
            return rankvit_forward(sd, cfg, images, 1.0 if budget is None else budget)
        if family == "residualvit":
            return residualvit_forward(sd, cfg, images, budget)
        if family == "adavit":
            return avit_forward(sd, cfg, images)
        if family == "moevit":
            return moevit_forward(sd, cfg, images)
        if family == "eeresidualvit":
            return eeresidualvit_forward(sd, cfg, images, budget)
    raise ValueError(family)


def flops_per_image(cfg, tokens_per_layer: Optional[Sequence[int]] = None) -> float:
    """Algorithmic FLOPs (2*MAC) per image, SURVEY.md §8d:
    ``2*P*Kp*D + sum_l(6 n D^2 + 4 n^2 D + 2 n D^2 + 4 n D F) + 2 D C``."""
    D, F_, L, C = cfg["hidden_dim"], cfg["mlp_dim"], cfg["num_layers"], cfg["num_classes"]
    P = (cfg["image_size"] // cfg["patch_size"]) ** 2
    Kp = 3 * cfg["patch_size"] ** 2
    if tokens_per_layer is None:
        tokens_per_layer = [P + cfg.get("num_class_tokens", 1) + cfg.get("num_registers", 0)] * L
    total = 2.0 * P * Kp * D + 2.0 * D * C
    for n in tokens_per_layer:
        total += 6.0 * n * D * D + 4.0 * n * n * D + 2.0 * n * D * D + 4.0 * n * D * F_
    return total
