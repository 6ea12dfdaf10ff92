"""Host-side orchestration of the encoder forward on the C-ABI kernels.

``pack_model`` reads the parameters of a module that follows the reference's attribute
layout (the mirrors in ``peekvit_b200.models`` *or* the reference's own classes) and keeps
GEMM operands as bf16 [N,K] and everything else as fp32.  ``Forward`` runs the layers on
packed token rows: the residual stream is fp32 ``[rows, D]``; LayerNorm writes bf16 GEMM
operands; every GEMM accumulates in fp32 (TMEM) with bias / GELU / residual fused.

Reference call stacks mirrored here: ``VisionTransformer.forward`` (models/vit.py:224-248),
``ViTEncoder.forward`` (vit.py:90-95), ``ViTBlock.forward`` (vit.py:45-55),
``RankViTBlock.sort_and_drop`` (rankvit.py:55-77).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Union

import torch

from . import ops
from ._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_F32, PK_EPI_BIAS_GELU_BF16, PK_EPI_BIAS_RESID_F32


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.bfloat16).contiguous()


@dataclass
class AttnWeights:
    w_qkv: torch.Tensor      # bf16 [3D, D]
    b_qkv: torch.Tensor      # f32 [3D]
    w_o: torch.Tensor        # bf16 [D, D]
    b_o: torch.Tensor        # f32 [D]
    bias_kv: torch.Tensor    # bf16 [2D]: what a zero token projects to (k-bias | v-bias)


@dataclass
class MlpWeights:
    w_fc1: torch.Tensor      # bf16 [F, D]
    b_fc1: torch.Tensor
    w_fc2: torch.Tensor      # bf16 [D, F]
    b_fc2: torch.Tensor


@dataclass
class LayerWeights:
    kind: str                                  # 'vit' | 'rank' | 'residual' | 'avit' | 'moe'
    ln1_w: torch.Tensor
    ln1_b: torch.Tensor
    ln2_w: torch.Tensor
    ln2_b: torch.Tensor
    eps: float
    attn: List[AttnWeights]
    mlp: List[MlpWeights]
    extra: Dict[str, Any] = field(default_factory=dict)
    module: Any = None                         # the live block (side-state is published on it)


@dataclass
class PackedModel:
    family: str
    image_size: int
    patch_size: int
    dim: int
    heads: int
    num_classes: int
    n_cls: int
    n_reg: int
    w_patch: torch.Tensor                      # bf16 [D, 3*p*p]
    b_patch: torch.Tensor
    cls_tokens: torch.Tensor                   # f32 [T, D]
    reg_tokens: Optional[torch.Tensor]
    pos: torch.Tensor                          # f32 [seq, D]
    layers: List[LayerWeights]
    ln_w: torch.Tensor
    ln_b: torch.Tensor
    ln_eps: float
    head_w: torch.Tensor                       # f32 [C, D]
    head_b: torch.Tensor
    extra: Dict[str, Any] = field(default_factory=dict)

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def seq_len(self) -> int:
        return self.num_patches + self.n_cls + self.n_reg


def _pack_attn(mha) -> AttnWeights:
    D = mha.in_proj_weight.shape[1]
    b = _f32(mha.in_proj_bias)
    return AttnWeights(_bf16(mha.in_proj_weight), b, _bf16(mha.out_proj.weight), _f32(mha.out_proj.bias),
                       _bf16(b[D:3 * D]))


def _pack_mlp(mlp) -> MlpWeights:
    return MlpWeights(_bf16(mlp.fc1.weight), _f32(mlp.fc1.bias), _bf16(mlp.fc2.weight), _f32(mlp.fc2.bias))


def _block_kind(blk) -> str:
    name = type(blk).__name__
    return {"ViTBlock": "vit", "RankViTBlock": "rank", "ResidualViTBlock": "residual", "AViTBlock": "avit",
            "ViTBlockMoE": "moe"}.get(name, name)


def pack_model(model, family: str) -> PackedModel:
    """Snapshot the live parameters of ``model`` (reference attribute layout)."""
    layers: List[LayerWeights] = []
    for blk in model.encoder.layers:
        kind = _block_kind(blk)
        if kind not in ("vit", "rank", "residual", "avit", "moe"):
            raise NotImplementedError(
                f"encoder.layers contains a {type(blk).__name__}; only the reference's transformer blocks run on the "
                "B200 path (NoiseBlock splicing, reference utils/utils.py:162-191, is outside the hot-path scope)")
        if kind == "moe":
            attn = [_pack_attn(e.self_attention) for e in blk.self_attention.experts]
            mlps = [_pack_mlp(e) for e in blk.mlp.experts]
            extra = {}
            if len(mlps) > 1:
                extra["mlp_gate_w"] = _f32(blk.mlp.gating_network.gate.weight)
                extra["mlp_gate_b"] = _f32(blk.mlp.gating_network.gate.bias)
            if len(attn) > 1:
                raise NotImplementedError("attention-MoE layers (reference moevit.py:71-102) are a 'next' row (SURVEY.md §8 f3)")
        else:
            attn = [_pack_attn(blk.self_attention.self_attention)]
            mlps = [_pack_mlp(blk.mlp)]
            extra = {}
        if kind == "residual":
            extra["skip"] = blk.skip
            if blk.skip in ("attention", "mlp"):
                raise NotImplementedError(f"ResidualViT skip mode {blk.skip!r} is a 'next' row (SURVEY.md §8 f3)")
            if blk.skip == "attention+mlp":
                g = blk.residual_gate
                extra.update(gate_w=_f32(g.projection.weight).reshape(-1), gate_b=_f32(g.projection.bias).reshape(-1),
                             gate_type=g.gate_type, gate_temp=float(g.temp), gate_bias=float(g.sigmoid_bias),
                             gate_threshold=g.threshold, budget_token=blk.budget_token, add_input=bool(blk.add_input))
                if blk.budget_token == "learnable":
                    extra["bt_gate_w"] = _f32(blk.budget_token_gate.weight).reshape(-1)
                    extra["bt_gate_b"] = _f32(blk.budget_token_gate.bias).reshape(-1)
        if kind == "avit":
            extra.update(gate_scale=float(blk.gate_scale), gate_center=float(blk.gate_center))
        layers.append(LayerWeights(kind, _f32(blk.ln_1.weight), _f32(blk.ln_1.bias), _f32(blk.ln_2.weight), _f32(blk.ln_2.bias),
                                   float(blk.ln_1.eps), attn, mlps, extra, blk))
    D = model.hidden_dim
    cls = model.class_token if family == "moevit" else model.class_tokens
    n_reg = int(getattr(model, "num_registers", 0) or 0) if family != "moevit" else 0
    pm = PackedModel(
        family=family, image_size=model.image_size, patch_size=model.patch_size, dim=D,
        heads=model.encoder.layers[0].num_heads if len(model.encoder.layers) else model.num_heads,
        num_classes=model.num_classes, n_cls=cls.shape[1], n_reg=n_reg,
        w_patch=_bf16(model.conv_proj.weight.reshape(D, -1)), b_patch=_f32(model.conv_proj.bias),
        cls_tokens=_f32(cls.reshape(-1, D)),
        reg_tokens=_f32(model.register_tokens.reshape(-1, D)) if n_reg > 0 else None,
        pos=_f32(model.encoder.pos_embedding.reshape(-1, D)),
        layers=layers, ln_w=_f32(model.encoder.ln.weight), ln_b=_f32(model.encoder.ln.bias), ln_eps=float(model.encoder.ln.eps),
        head_w=_f32(model.head.weight), head_b=_f32(model.head.bias))
    return pm


def params_fingerprint(model) -> tuple:
    """Changes whenever a parameter is rebound, moved or modified in place."""
    return tuple((p.data_ptr(), p._version) for p in model.parameters()) + (len(model.encoder.layers),)


class Workspace:
    """Torch-allocated scratch buffers, reused across calls (stable addresses keep the TMA
    descriptor cache warm and make the launch sequence CUDA-graph capturable)."""

    def __init__(self, device):
        self.device = device
        self._bufs: Dict[tuple, torch.Tensor] = {}

    def get(self, name: str, shape: Sequence[int], dtype) -> torch.Tensor:
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t


class Forward:
    """One forward pass over a micro-batch of images already resident on the device."""

    def __init__(self, pm: PackedModel, ws: Workspace):
        self.pm, self.ws = pm, ws

    # ---------------------------------------------------------------- shared pieces
    def embed(self, images: torch.Tensor, extra_rows: int = 0) -> torch.Tensor:
        """Patch GEMM (+conv bias +pos_embedding) and class/register rows -> x f32 [B*seq, D]
        (reference vit.py:203-236, :92).  ``extra_rows`` reserves trailing rows per sample."""
        pm = self.pm
        B = images.shape[0]
        P, D, T, R = pm.num_patches, pm.dim, pm.n_cls, pm.n_reg
        seq = pm.seq_len + extra_rows
        patches = ops.patchify(images, pm.patch_size, self.ws.get("patches", (B * P, pm.w_patch.shape[1]), torch.bfloat16))
        x = self.ws.get("x", (B * seq, D), torch.float32)
        ops.gemm(patches, pm.w_patch, pm.b_patch, x, PK_EPI_BIAS_RESID_F32, resid=pm.pos,
                 rows_per_group=P, group_stride=seq, group_offset=T + R, resid_is_pos=True)
        ops.fill_token_rows(x, B, seq, 0, pm.cls_tokens, pm.pos)
        if R > 0:
            ops.fill_token_rows(x, B, seq, T, pm.reg_tokens, pm.pos)
        return x

    def dense_block(self, x: torch.Tensor, lw: LayerWeights, rows: int, batch: int, seq: int) -> None:
        """ViTBlock on uniform-length samples, in place on the fp32 residual stream
        (reference vit.py:45-55)."""
        pm, ws = self.pm, self.ws
        D = pm.dim
        aw, mw = lw.attn[0], lw.mlp[0]
        F = mw.w_fc1.shape[0]
        a = ops.layernorm(x, lw.ln1_w, lw.ln1_b, lw.eps, ws.get("ln", (rows, D), torch.bfloat16), rows=rows)
        qkv = ops.gemm(a, aw.w_qkv, aw.b_qkv, ws.get("qkv", (rows, 3 * D), torch.bfloat16), PK_EPI_BIAS_BF16)
        att = ops.attention(qkv, ws.get("att", (rows, D), torch.bfloat16), batch, pm.heads, D // pm.heads, seq_len=seq)
        ops.gemm(att, aw.w_o, aw.b_o, x[:rows], PK_EPI_BIAS_RESID_F32, resid=x[:rows])
        a = ops.layernorm(x, lw.ln2_w, lw.ln2_b, lw.eps, ws.get("ln", (rows, D), torch.bfloat16), rows=rows)
        hid = ops.gemm(a, mw.w_fc1, mw.b_fc1, ws.get("hid", (rows, F), torch.bfloat16), PK_EPI_BIAS_GELU_BF16)
        ops.gemm(hid, mw.w_fc2, mw.b_fc2, x[:rows], PK_EPI_BIAS_RESID_F32, resid=x[:rows])

    def head(self, x: torch.Tensor, batch: int, seq: int, cu_seqlens=None) -> torch.Tensor:
        pm = self.pm
        return ops.cls_head(x, batch, seq, pm.n_cls, pm.ln_w, pm.ln_b, pm.ln_eps, pm.head_w, pm.head_b,
                            cu_seqlens=cu_seqlens, out=self.ws.get("logits", (batch, pm.num_classes), torch.float32))

    # ---------------------------------------------------------------- plain ViT
    def vit(self, images: torch.Tensor) -> torch.Tensor:
        pm = self.pm
        B, seq = images.shape[0], pm.seq_len
        x = self.embed(images)
        for lw in pm.layers:
            self.dense_block(x, lw, B * seq, B, seq)
        return self.head(x, B, seq)

    # ---------------------------------------------------------------- RankViT
    def rankvit(self, images: torch.Tensor, budgets: Dict[int, float], aux: Optional[dict] = None) -> torch.Tensor:
        """``budgets`` maps layer index -> current_budget of the RankViTBlocks (reference
        rankvit.py:283-288).  A rank layer with budget != 1 scores the non-class tokens of its
        input by L2 norm, keeps the top ceil(n*b) in descending order (ties -> lowest index)
        and compacts them behind the class token (rankvit.py:55-77); later layers run on the
        survivors only."""
        pm, ws = self.pm, self.ws
        B, seq, D = images.shape[0], pm.seq_len, pm.dim
        x = self.embed(images)
        flip = 0
        for i, lw in enumerate(pm.layers):
            b = budgets.get(i, 1.0) if lw.kind == "rank" else 1.0
            if lw.kind == "rank" and b != 1:
                n = seq - 1
                k = math.ceil(n * b)
                scores = ops.token_norm_score(x, B, seq, ws.get(f"rank_scores_{n}", (B, n), torch.float32))
                kept = ops.topk_select(scores, k, ws.get(f"rank_kept_{n}_{k}", (B, k), torch.int32))
                y = ws.get(f"x_compact_{flip}", (B * (k + 1), D), torch.float32)
                flip ^= 1
                ops.gather_rows(x, kept, B, seq, y)
                if aux is not None:
                    aux.setdefault("scores", {})[i] = scores
                    aux.setdefault("kept", {})[i] = kept
                x, seq = y, k + 1
            if aux is not None:
                aux.setdefault("seq_lens", []).append(seq)
            self.dense_block(x, lw, B * seq, B, seq)
        return self.head(x, B, seq)
