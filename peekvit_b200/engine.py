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

import os

from . import ops
from ._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_F32, PK_EPI_BIAS_GELU_BF16, PK_EPI_BIAS_RESID_F32, PK_OUT_BF16X2, PK_OUT_F16
from .ops import SPLIT_GELU, SPLIT_LAYERNORM, SPLIT_NONE


# LayerNorm folded across GEMMs: 0 = separate LayerNorm kernels everywhere; 1 = ln_1 folded into the in-projection (the
# previous layer's fc2 epilogue emits the bf16 row copy + statistics), ln_2 a kernel; 2 = both folded.  Measured on ViT-B/16
# (B200, 256-image micro-batch): the GELU epilogue of fc1 is on the critical path of that GEMM, so folding ln_2 into it costs
# more (+41 us) than the LayerNorm kernel it removes (42 us) -> default 1.
LN_FOLD = int(os.environ.get("PEEKVIT_B200_LN_FOLD", "1"))
# 1: the im2col operand has one row per token, so the patch GEMM runs on the CTA-pair kernel and writes the residual
# stream (and layer 0's LayerNorm statistics) directly; 0: densely packed patches + row-remap epilogue (single-CTA kernel)
EMBED_TOKEN_ROWS = int(os.environ.get("PEEKVIT_B200_EMBED_TOKEN_ROWS", "1"))
# 1: for more than 256 samples the classification head runs as a split-operand tensor-core GEMM; 0: always the fused kernel
HEAD_GEMM = int(os.environ.get("PEEKVIT_B200_HEAD_GEMM", "1"))
# MoE expert fc2: un-permute + residual add inside the GEMM epilogue (row-indexed reductions) instead of a sorted fp32 output
# followed by pk_scatter_add_rows (A/B switch)
# ragged attention: mean live rows per sample from which the two-region tcgen05 kernel beats the general mma.sync kernel
# (tools/attn_ragged_bench.py, profiles/r02)
ATT_TCR_MIN_MEAN_ROWS = int(os.environ.get("PEEKVIT_B200_ATT_TCR_MIN_MEAN_ROWS", "140"))
MOE_FUSED_SCATTER = int(os.environ.get("PEEKVIT_B200_MOE_FUSED_SCATTER", "1"))
# ... and one grouped launch per projection over all expert segments instead of one launch per expert (A/B switch)
MOE_GROUPED = int(os.environ.get("PEEKVIT_B200_MOE_GROUPED", "1"))


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


def _pack_ln_fold(blk) -> Dict[str, torch.Tensor]:
    """Operands of the LayerNorm-fused GEMM chain (include/peekvit_b200.h, pk_gemm_args.ln_*):
    LN(x) @ W^T + b  ==  rstd * (x @ (gamma*W)^T) - rstd*mean * c1 + c2  with  c1 = sum_k (gamma*W)[n,k],
    c2 = b + W @ beta.  The gain is folded into the bf16 weight from the fp32 master; c1 sums the *rounded*
    folded weight so the mean term cancels exactly what the tensor core accumulates."""
    def fold(w, b, g, be):
        w32, g32, be32 = w.detach().float(), g.detach().float(), be.detach().float()
        wf = (w32 * g32[None, :]).to(torch.bfloat16).contiguous()
        return wf, wf.float().sum(1).contiguous(), (b.detach().float() + w32 @ be32).contiguous()
    mha = blk.self_attention.self_attention
    w_qkv, c1_qkv, c2_qkv = fold(mha.in_proj_weight, mha.in_proj_bias, blk.ln_1.weight, blk.ln_1.bias)
    w_fc1, c1_fc1, c2_fc1 = fold(blk.mlp.fc1.weight, blk.mlp.fc1.bias, blk.ln_2.weight, blk.ln_2.bias)
    return dict(w_qkv=w_qkv, c1_qkv=c1_qkv, c2_qkv=c2_qkv, w_fc1=w_fc1, c1_fc1=c1_fc1, c2_fc1=c2_fc1)


def _block_kind(blk) -> str:
    name = type(blk).__name__
    return {"ViTBlock": "vit", "RankViTBlock": "rank", "ResidualViTBlock": "residual", "AViTBlock": "avit",
            "ViTBlockMoE": "moe", "NoiseBlock": "noise"}.get(name, name)


def draw_noise(pm: "PackedModel", batch: int, device, budgets: Optional[Dict[int, float]] = None) -> Dict[int, tuple]:
    """The random draws of every NoiseBlock for ONE forward of the whole batch, in layer order, from torch's generators like
    the reference (``randn_like`` on the device generator, ``randperm`` on the host generator; blocks.py:117-157): the same
    token positions are dropped in every sample of the batch whatever the micro-batch size, the Gaussian draw is one
    (B, N, D) tensor, and it is made even at 0 dB (blocks.py:129) so that the generator state stays aligned with the reference."""
    draws: Dict[int, tuple] = {}
    seq = pm.seq_len
    # ResidualViT: the reference's budget token is the LAST token of the sequence (residualvit.py:572-583); the dense row layout
    # keeps it at local row 1 (Forward.residualvit_dense), so draws made in the reference's token order are permuted
    budget_last = pm.family == "residualvit" and bool(pm.extra.get("add_budget_token"))
    if budget_last:
        seq += 1
    for i, lw in enumerate(pm.layers):
        if lw.kind == "rank" and budgets is not None:
            b = budgets.get(i, 1.0)
            if b != 1:
                seq = min(max(math.ceil((seq - 1) * b), 0), seq - 1) + 1
        if lw.kind == "noise":
            kind, val, db = draw_block_noise(lw.module, batch, seq, pm.dim, device)
            if budget_last and val is not None:
                if kind == "snr":
                    order = torch.tensor([0, seq - 1] + list(range(1, seq - 1)), device=device)     # layout row -> reference token
                    val = val[:, order].contiguous()
                else:
                    v = val.long()
                    val = torch.where(v == 0, v, torch.where(v == seq - 1, torch.ones_like(v), v + 1)).to(torch.int32)
            draws[i] = (kind, val, db)
    return draws


def draw_block_noise(blk, batch: int, seq: int, dim: int, device) -> tuple:
    """One NoiseBlock's draw for a (batch, seq, dim) input: ("snr", noise or None, dB) / ("drop", token indices or None, 0)."""
    if blk.snr_db is not None:
        noise = torch.randn((batch, seq, dim), dtype=torch.float32, device=device)
        return ("snr", noise if blk.snr_db != 0 else None, float(blk.snr_db))
    if blk.std is not None:
        raise ValueError("std is not supported anymore. Please use snr instead.")
    idx = None
    if blk.prob != 0:
        num_mask = int(blk.prob * seq)           # blocks.py:151 (TypeError for an unset block, like the reference)
        perm = torch.randperm(seq)[:num_mask]
        idx = perm.to(device=device, dtype=torch.int32) if num_mask > 0 else None
    return ("drop", idx, 0.0)


def apply_noise(draw: tuple, x: torch.Tensor, batch: int, seq: int, offset: int = 0) -> None:
    """NoiseBlock.forward (reference blocks.py:159-170) on the fp32 token rows x [batch*seq, D] of the samples
    [offset, offset + batch) of the forward ``draw`` was made for (``draw_noise``), in place: Gaussian noise scaled per token
    to the requested SNR, or the same token positions zeroed in every sample."""
    kind, val, snr_db = draw
    if val is None:
        return
    if kind == "snr":
        if val.shape[1] != seq:
            raise RuntimeError(f"NoiseBlock draw was made for {val.shape[1]} tokens per sample, the block sees {seq}")
        ops.noise_snr(x, val[offset:offset + batch].reshape(batch * seq, x.shape[-1]), snr_db, batch * seq)
    else:
        ops.zero_token_rows(x, batch, seq, val)


def pack_model(model, family: str) -> PackedModel:
    """Snapshot the live parameters of ``model`` (reference attribute layout)."""
    layers: List[LayerWeights] = []
    for blk in model.encoder.layers:
        kind = _block_kind(blk)
        if kind == "noise":
            if family not in ("vit", "rankvit", "moevit", "residualvit", "eeresidualvit"):
                # the compacted A-ViT row layout does not materialise halted rows, which independent per-row noise would make
                # distinct -- and the reference's A-ViT loop cannot run a NoiseBlock either (adavit.py:170-176).  A ResidualViT
                # encoder with a NoiseBlock runs on the dense (masked, uncompacted) row layout instead.
                raise NotImplementedError(f"NoiseBlock inside a {family} encoder is not supported on the B200 path")
            layers.append(LayerWeights("noise", None, None, None, None, 0.0, [], [], {}, blk))
            continue
        if kind not in ("vit", "rank", "residual", "avit", "moe"):
            raise NotImplementedError(
                f"encoder.layers contains a {type(blk).__name__}; only the reference's transformer blocks and NoiseBlock "
                "run on the B200 path")
        if kind == "moe":
            attn = [_pack_attn(e.self_attention) for e in blk.self_attention.experts]
            mlps = [_pack_mlp(e) for e in blk.mlp.experts]
            extra = {}
            if len(mlps) > 1:
                extra["mlp_gate_w"] = _f32(blk.mlp.gating_network.gate.weight)
                extra["mlp_gate_b"] = _f32(blk.mlp.gating_network.gate.bias)
            if len(attn) > 1:
                extra["attn_gate_w"] = _f32(blk.self_attention.gating_network.gate.weight)
                extra["attn_gate_b"] = _f32(blk.self_attention.gating_network.gate.bias)
        else:
            attn = [_pack_attn(blk.self_attention.self_attention)]
            mlps = [_pack_mlp(blk.mlp)]
            extra = {}
        if kind == "residual":
            extra["skip"] = blk.skip
            if blk.skip in ("attention", "mlp", "attention+mlp"):
                g = blk.residual_gate
                extra.update(gate_w=_f32(g.projection.weight).reshape(-1), gate_b=float(g.projection.bias.detach().float().cpu()[0]),
                             gate_type=g.gate_type, gate_temp=float(g.temp), gate_bias=float(g.sigmoid_bias),
                             gate_threshold=g.threshold, budget_token=blk.budget_token, add_input=bool(blk.add_input))
                if blk.budget_token == "learnable" and blk.skip == "attention+mlp":     # the one mode that reads it (:212)
                    extra["bt_gate_w"] = _f32(blk.budget_token_gate.weight).reshape(-1)
                    extra["bt_gate_b"] = float(blk.budget_token_gate.bias.detach().float().cpu()[0])
        if kind == "avit":
            extra.update(gate_scale=float(blk.gate_scale), gate_center=float(blk.gate_center))
        if kind in ("vit", "rank"):
            extra["fold"] = _pack_ln_fold(blk)
        layers.append(LayerWeights(kind, _f32(blk.ln_1.weight), _f32(blk.ln_1.bias), _f32(blk.ln_2.weight), _f32(blk.ln_2.bias),
                                   float(blk.ln_1.eps), attn, mlps, extra, blk))
    D = model.hidden_dim
    cls = model.class_token if family == "moevit" else model.class_tokens
    n_reg = int(getattr(model, "num_registers", 0) or 0) if family != "moevit" else 0
    pm = PackedModel(
        family=family, image_size=model.image_size, patch_size=model.patch_size, dim=D,
        heads=next((b.num_heads for b in model.encoder.layers if hasattr(b, "num_heads")), model.num_heads),
        num_classes=model.num_classes, n_cls=cls.shape[1], n_reg=n_reg,
        w_patch=_bf16(model.conv_proj.weight.reshape(D, -1)), b_patch=_f32(model.conv_proj.bias),
        cls_tokens=_f32(cls.reshape(-1, D)),
        reg_tokens=_f32(model.register_tokens.reshape(-1, D)) if n_reg > 0 else None,
        pos=_f32(model.encoder.pos_embedding.reshape(-1, D)),
        layers=layers, ln_w=_f32(model.encoder.ln.weight), ln_b=_f32(model.encoder.ln.bias), ln_eps=float(model.encoder.ln.eps),
        head_w=_f32(model.head.weight), head_b=_f32(model.head.bias))
    pm.extra["conv_weight"] = model.conv_proj.weight          # fp32 master, read lazily by the fp32-accurate mode
    if family == "eeresidualvit":
        # EEResidualVisionTransformer (eeresidualvit.py): the ResidualViT forward + one LayerNorm -> Linear head per layer on
        # the first class token (:73-75,:91-96); its budget mode lives in ``.budget`` (:170)
        family = pm.family = "residualvit"
        pm.extra["ee_heads"] = [(_f32(h[0].weight), _f32(h[0].bias), float(h[0].eps), _f32(h[1].weight), _f32(h[1].bias))
                                for h in model.encoder.early_exit_heads]
        pm.extra["add_budget_token"] = model.budget
    elif family == "residualvit":
        pm.extra["add_budget_token"] = model.add_budget_token
    if family == "residualvit":
        model_abt = pm.extra["add_budget_token"]
        gated = [lw for lw in pm.layers if lw.extra.get("skip") in ("attention", "mlp", "attention+mlp")]
        # Compacted rows (dropped tokens leave the batch) need every gated layer to be the shipped 'attention+mlp' block with
        # one class token and no add_input; everything else the reference can run -- the 'attention' / 'mlp' skip modes
        # (residualvit.py:130-194), add_input (:189-192), several class tokens, a NoiseBlock between the blocks -- makes
        # dropped rows distinct again (or readable by the head) and runs on the dense masked layout (Forward.residualvit_dense).
        pm.extra["dense_layout"] = bool(
            any(lw.kind == "noise" for lw in pm.layers)
            or any(lw.extra["skip"] != "attention+mlp" or lw.extra["add_input"] for lw in gated)
            or (gated and pm.n_cls != 1))            # further class tokens are gated like image tokens but read by the head
        if model_abt in ("learnable", "learnable_interpolate"):
            pm.extra["budget_token_1"] = _f32(model.learnable_budget_token_1.reshape(1, D))
        if model_abt == "learnable_interpolate":
            pm.extra["budget_token_2"] = _f32(model.learnable_budget_token_2.reshape(1, D))
        for lw in pm.layers:
            if lw.extra.get("skip") == "attention+mlp" and not pm.extra["dense_layout"]:
                blk = lw.module
                # what a dropped (zero) row equals when it leaves the block: fc2(gelu(fc1.bias)) + fc2.bias
                h = torch.nn.functional.gelu(blk.mlp.fc1.bias.detach().float())
                lw.extra["mlp0"] = (torch.nn.functional.linear(h, blk.mlp.fc2.weight.detach().float(),
                                                               blk.mlp.fc2.bias.detach().float())).contiguous()
    if family == "adavit":
        if pm.n_cls != 1:
            raise NotImplementedError("the B200 AViT path supports num_class_tokens == 1 (all shipped configs)")
        pm.extra["eps"] = float(model.eps)
    return pm


LIGHT_PARAMS = ("class_tokens", "class_token", "head.weight", "head.bias")
_LIGHT_GATE_SUFFIXES = (".residual_gate.projection.weight", ".residual_gate.projection.bias", ".budget_token_gate.weight",
                        ".budget_token_gate.bias")


def is_light_param(name: str) -> bool:
    """Parameters of the fine-tuning regimes (train/train.py:99-100) that an optimiser step changes while the backbone stays
    frozen: class tokens and head; for ResidualViT also the gate projections, budget-token gates and learnable budget tokens."""
    return (name in LIGHT_PARAMS or name.startswith("learnable_budget_token_")
            or (name.startswith("encoder.layers.") and name.endswith(_LIGHT_GATE_SUFFIXES)))


def _refresh(dst: torch.Tensor, src: torch.Tensor) -> None:
    """dst <- src unless dst already IS the live parameter's storage (fp32 parameters are packed without a copy; an in-place
    copy onto itself would bump the parameter's version counter and make every later call look like another update)."""
    src = src.detach()
    if dst.data_ptr() != src.data_ptr():
        dst.copy_(src.reshape(dst.shape))


def refresh_light(pm: "PackedModel", model) -> None:
    """Re-read the light parameters from the live module into an existing weight pack (same storage), and drop what was
    derived from them: the embedding's initial rows and the split head weight."""
    cls = model.class_token if pm.family == "moevit" else model.class_tokens
    _refresh(pm.cls_tokens, cls)
    _refresh(pm.head_w, model.head.weight)
    _refresh(pm.head_b, model.head.bias)
    pm.__dict__.pop("_embed_consts", None)
    pm.extra.pop("head_w6", None)
    if pm.family == "residualvit":
        for k, attr in (("budget_token_1", "learnable_budget_token_1"), ("budget_token_2", "learnable_budget_token_2")):
            if k in pm.extra:
                _refresh(pm.extra[k], getattr(model, attr))
        for lw in pm.layers:
            if "gate_w" in lw.extra:
                g = lw.module.residual_gate
                _refresh(lw.extra["gate_w"], g.projection.weight)
                lw.extra["gate_b"] = float(g.projection.bias.detach().float().cpu()[0])
            if "bt_gate_w" in lw.extra:
                _refresh(lw.extra["bt_gate_w"], lw.module.budget_token_gate.weight)
                lw.extra["bt_gate_b"] = float(lw.module.budget_token_gate.bias.detach().float().cpu()[0])


def params_fingerprint(model) -> tuple:
    """Changes whenever a parameter is rebound, moved or modified in place, or ``encoder.layers`` is edited
    (layers deleted, a parameter-free NoiseBlock spliced in)."""
    return tuple((p.data_ptr(), p._version) for p in model.parameters()) + tuple(id(b) for b in model.encoder.layers)


class Workspace:
    """Torch-allocated scratch buffers, reused across calls (stable addresses keep the TMA
    descriptor cache warm and make the launch sequence CUDA-graph capturable)."""

    def __init__(self, device):
        self.device = device
        self._bufs: Dict[tuple, torch.Tensor] = {}

    def get(self, name: str, shape: Sequence[int], dtype, zero: bool = False) -> torch.Tensor:
        """``zero``: cleared when first allocated (buffers whose rows past the live ones are read -- and masked -- by a kernel
        must hold finite values there: 0 x NaN is NaN)."""
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t


class Forward:
    """One forward pass over a micro-batch of images already resident on the device."""

    def __init__(self, pm: PackedModel, ws: Workspace):
        self.pm, self.ws = pm, ws
        # set by the CUDA-graph runner: the im2col of the current micro-batch is already in the workspace, so the captured
        # sequence does not depend on where the caller's image tensor lives
        self.patches_ready = False
        self.input_norm = (ops.IMAGENET_MEAN, ops.IMAGENET_STD)      # Normalize statistics of the uint8 input path
        # Split-operand arithmetic modes (runner: model.pk_precision; csrc/pk_exact.cu).  terms = 3: "fp32", three bf16 terms
        # per operand (K' = 6K) + fp32 attention: logits within 1e-5.  terms = 2: "bf16x2", hi + lo (K' = 3K, the activation
        # row stored once as [lo | hi]) with the split produced by the GEMM / attention epilogues themselves and the tcgen05
        # attention on IEEE-half operands: 16 significant bits on every linear operand, 11 in the attention core.
        self.terms = 0
        # NoiseBlock draws of the current forward (draw_noise) and the first sample of the micro-batch being run
        self.noise_draws: Dict[int, tuple] = {}
        self.sample_offset = 0

    @property
    def exact(self) -> bool:
        return self.terms > 0

    def _wrap(self, k: int) -> dict:
        """gemm() keyword for a split activation operand of unsplit width k."""
        return dict(a_wrap_k=k, cta_pair=2) if self.terms == 2 else {}

    def _embed_geometry(self, batch: int):
        """(tokens per sample of the embedded stream, shift rows after the first class token, token-row layout?).
        With the token-row layout the im2col operand has one row per TOKEN (class / register / budget rows zero), so the
        patch GEMM writes the residual stream in place of a row remap and runs on the CTA-pair kernel."""
        pm = self.pm
        shift = 1 if (pm.family == "residualvit" and pm.extra.get("add_budget_token")) else 0
        seq = pm.seq_len + shift
        token_rows = EMBED_TOKEN_ROWS and batch * seq > 256 and pm.dim % 8 == 0
        return seq, shift, token_rows

    def _patch_operand(self, batch: int) -> torch.Tensor:
        pm = self.pm
        seq, _, token_rows = self._embed_geometry(batch)
        Kp = pm.w_patch.shape[1]
        if not token_rows:
            return self.ws.get("patches", (batch * pm.num_patches, Kp), torch.bfloat16)
        key = ("patch_rows", batch, seq, Kp)
        t = self.ws._bufs.get(key)
        if t is None:                               # zeroed once: the non-patch rows are never written again
            t = self.ws._bufs[key] = torch.zeros((batch * seq, Kp), dtype=torch.bfloat16, device=self.ws.device)
        return t

    def patchify(self, images: torch.Tensor) -> torch.Tensor:
        """im2col of float NCHW images, or of uint8 NHWC images with ToTensor + Normalize fused in; in the
        split-operand modes straight into split rows.  The only kernel that reads the caller's image tensor: the CUDA-graph
        runner launches it eagerly in front of every replay."""
        pm = self.pm
        B = images.shape[0]
        if self.exact:
            if images.dtype != torch.float32:
                raise NotImplementedError("the split-operand modes take float images (the uint8 input path is a bf16-mode feature)")
            t, Kp = self.terms, pm.w_patch.shape[1]
            return ops.patchify_split(images, pm.patch_size,
                                      self.ws.get(f"patches_x{t}", (B * pm.num_patches, (6 if t == 3 else 3) * Kp), torch.bfloat16), t)
        seq, shift, token_rows = self._embed_geometry(B)
        out = self._patch_operand(B)
        lay = dict(rows_per_sample=seq, row_offset=pm.n_cls + pm.n_reg + shift) if token_rows else {}
        if images.dtype == torch.uint8:
            return ops.patchify_u8(images, pm.patch_size, out, *self.input_norm, **lay)
        return ops.patchify(images, pm.patch_size, out, **lay)

    # ---------------------------------------------------------------- shared pieces
    def embed(self, images: torch.Tensor, shift: int = 0, emit_fold: Optional[bool] = None):
        """Patch GEMM (+conv bias +pos_embedding) and class/register rows -> x f32 [B*seq, D]
        (reference vit.py:203-236, :92).  ``shift`` local rows are left free right after the first
        class token (ResidualViT budget token).  ``emit_fold`` not None: return (x, fold) where fold is the bf16 copy and
        row statistics of x for a LayerNorm-folded first layer, or None when not asked for / the patch GEMM cannot emit them."""
        pm = self.pm
        B = images.shape[0]
        P, D, T, R = pm.num_patches, pm.dim, pm.n_cls, pm.n_reg
        seq, shift_g, token_rows = self._embed_geometry(B)
        assert shift == shift_g
        patches = self._patch_operand(B) if self.patches_ready else self.patchify(images)
        x = self.ws.get("x", (B * seq, D), torch.float32)
        if token_rows:
            dev = x.device

            def build_init():
                # what every sample's rows hold before the patch projection is added: class / register tokens + their
                # positions, the position embedding alone on patch rows, zeros on the budget-token row
                t = torch.zeros((B * seq, D), dtype=torch.float32, device=dev)
                ops.fill_token_rows(t, B, seq, 0, pm.cls_tokens[:1], pm.pos)
                if T > 1:
                    ops.fill_token_rows(t, B, seq, 1 + shift, pm.cls_tokens[1:], pm.pos, pos_offset=1)
                if R > 0:
                    ops.fill_token_rows(t, B, seq, T + shift, pm.reg_tokens, pm.pos, pos_offset=T)
                ops.fill_token_rows(t, B, seq, T + R + shift, None, pm.pos, pos_offset=T + R, n_tokens=P, scale=0.0)
                return t

            def build_scale():
                rs = torch.zeros((B, seq), dtype=torch.float32, device=dev)
                rs[:, T + R + shift:T + R + shift + P] = 1.0
                return rs.reshape(-1)

            consts = pm.__dict__.setdefault("_embed_consts", {})          # lives and dies with the weight pack
            x_init = consts.get((B, seq))
            if x_init is None:
                x_init = consts[(B, seq)] = build_init()
            rs = self._const(f"patch_rowscale_{B}_{seq}_{T}_{R}", build_scale)
            fold = None
            if emit_fold:
                fold = (self.ws.get("xb", (B * seq, D), torch.bfloat16),
                        self.ws.get("ln_stats", (B * seq, ops.gemm_row_stat_parts(D), 2), torch.float32))
                ops.gemm(patches, pm.w_patch, pm.b_patch, x, PK_EPI_BIAS_RESID_F32, resid=x_init, rowscale=rs,
                         xb_out=fold[0], row_stats=fold[1])
            else:
                ops.gemm(patches, pm.w_patch, pm.b_patch, x, PK_EPI_BIAS_RESID_F32, resid=x_init, rowscale=rs)
            return x if emit_fold is None else (x, fold)
        ops.gemm(patches, pm.w_patch, pm.b_patch, x, PK_EPI_BIAS_RESID_F32, resid=pm.pos,
                 rows_per_group=P, group_stride=seq, group_offset=T + R + shift, resid_is_pos=True, pos_offset=T + R)
        ops.fill_token_rows(x, B, seq, 0, pm.cls_tokens[:1], pm.pos)
        if T > 1:
            ops.fill_token_rows(x, B, seq, 1 + shift, pm.cls_tokens[1:], pm.pos, pos_offset=1)
        if R > 0:
            ops.fill_token_rows(x, B, seq, T + shift, pm.reg_tokens, pm.pos, pos_offset=T)
        return x if emit_fold is None else (x, None)

    def attn_part(self, x, lw: LayerWeights, rows: int, batch: int, *, seq: int = 0, cu=None, max_len: int = 0, rows_dev=None,
                  rowscale=None, key_mult=None, extra_mult=None) -> None:
        """x += rowscale * out_proj(attention(in_proj(rowscale * LN1(x))))  (vit.py:48-51; residualvit.py:252-256)."""
        if self.exact:
            return self._attn_part_exact(x, lw, rows, batch, seq, cu, max_len, rows_dev, rowscale, key_mult, extra_mult)
        pm, ws = self.pm, self.ws
        D = pm.dim
        aw = lw.attn[0]
        a = ops.layernorm(x, lw.ln1_w, lw.ln1_b, lw.eps, ws.get("ln", (rows, D), torch.bfloat16), rows=rows, rowscale=rowscale,
                          rows_dev=rows_dev)
        # zeroed once: the ragged tcgen05 attention loads fixed-size key tiles, i.e. also rows past the live ones (masked)
        qkv = ops.gemm(a, aw.w_qkv, aw.b_qkv, ws.get("qkv", (rows, 3 * D), torch.bfloat16, zero=True), PK_EPI_BIAS_BF16, m_dev=rows_dev)
        # ragged batches: the quad-region tcgen05 kernel takes the samples of <= 128 keys; the longer ones go to the two-region
        # tcgen05 kernel when the live rows average ATT_TCR_MIN_MEAN_ROWS per sample or more, to the general mma.sync kernel
        # below that -- all decided on the device from cu_seqlens / the live row count (every launch is in the graph)
        att = ops.attention(qkv, ws.get("att", (rows, D), torch.bfloat16), batch, pm.heads, D // pm.heads, seq_len=seq,
                            cu_seqlens=cu, max_seq_len=max_len, key_mult=key_mult,
                            extra_kv=aw.bias_kv if extra_mult is not None else None, extra_mult=extra_mult,
                            route_rows=rows_dev if cu is not None else None, route_min_rows=batch * ATT_TCR_MIN_MEAN_ROWS)
        ops.gemm(att, aw.w_o, aw.b_o, x[:rows], PK_EPI_BIAS_RESID_F32, resid=x[:rows], rowscale=rowscale, m_dev=rows_dev)

    def mlp_part(self, x, lw: LayerWeights, rows: int, *, rows_dev=None, rowscale=None, residual: bool = True) -> None:
        """x += fc2(gelu(fc1(rowscale * LN2(x))))  (vit.py:53-55; residualvit.py:258-260); ``residual=False``: x = fc2(...) alone,
        what forward_skip_attention / forward_skip_mlp return (residualvit.py:153-157,186-187)."""
        if self.exact:
            return self._mlp_part_exact(x, lw, rows, rows_dev, rowscale, residual)
        pm, ws = self.pm, self.ws
        D = pm.dim
        mw = lw.mlp[0]
        F = mw.w_fc1.shape[0]
        a = ops.layernorm(x, lw.ln2_w, lw.ln2_b, lw.eps, ws.get("ln", (rows, D), torch.bfloat16), rows=rows, rowscale=rowscale,
                          rows_dev=rows_dev)
        hid = ops.gemm(a, mw.w_fc1, mw.b_fc1, ws.get("hid", (rows, F), torch.bfloat16), PK_EPI_BIAS_GELU_BF16, m_dev=rows_dev)
        if residual:
            ops.gemm(hid, mw.w_fc2, mw.b_fc2, x[:rows], PK_EPI_BIAS_RESID_F32, resid=x[:rows], m_dev=rows_dev)
        else:
            ops.gemm(hid, mw.w_fc2, mw.b_fc2, x[:rows], PK_EPI_BIAS_F32, m_dev=rows_dev)

    def dense_block(self, x: torch.Tensor, lw: LayerWeights, rows: int, batch: int, seq: int) -> None:
        """ViTBlock on uniform-length samples, in place on the fp32 residual stream (vit.py:45-55)."""
        self.attn_part(x, lw, rows, batch, seq=seq)
        self.mlp_part(x, lw, rows)

    # ---- fp32-accurate mode (the reference's shipped dtype): same tcgen05 GEMM kernels on 3-way split bf16 operands
    def _exact_weights(self, lw: LayerWeights) -> Dict[str, List[torch.Tensor]]:
        """Split weight rows of a block ([m|h|l|h|m|h] for three terms, [h|l|h] for two; one entry per expert, plain blocks
        have one), built from the fp32 masters on first use."""
        key = f"x{self.terms}"
        c = lw.extra.get(key)
        if c is None:
            blk, t = lw.module, self.terms
            if lw.kind == "moe":
                mhas = [e.self_attention for e in blk.self_attention.experts]
                mlps = list(blk.mlp.experts)
            else:
                mhas, mlps = [blk.self_attention.self_attention], [blk.mlp]
            c = lw.extra[key] = dict(w_qkv=[ops.split_weight(m.in_proj_weight, t) for m in mhas],
                                     w_o=[ops.split_weight(m.out_proj.weight, t) for m in mhas],
                                     w_fc1=[ops.split_weight(m.fc1.weight, t) for m in mlps],
                                     w_fc2=[ops.split_weight(m.fc2.weight, t) for m in mlps])
        return c

    def gemm_exact(self, a32: torch.Tensor, rows: int, w6: torch.Tensor, bias, out, epilogue: int, mode: int = SPLIT_NONE,
                   gamma=None, beta=None, eps: float = 0.0, resid=None, *, in_scale=None, row_index=None, rows_dev=None,
                   a6=None, **gemm_kw):
        """out = epilogue(in_scale * pre(a32) @ W^T + bias) at split-operand accuracy: pre = none / exact GELU / LayerNorm is
        applied by the kernel that splits the activation rows, the product runs as one bf16 GEMM over K' = 6K (three terms) or
        3K (two terms, rows stored as [lo | hi]).  ``a6``: reuse rows that are already split (several experts reading the same
        activations)."""
        K = a32.shape[-1]
        if a6 is None:
            a6 = self.ws.get(f"a{self.terms}_{K}", (rows, ops.split_width(self.terms) * K), torch.bfloat16)
            ops.split(a32, a6, self.terms, mode, gamma, beta, eps, rows=rows, rowscale=in_scale, row_index=row_index, rows_dev=rows_dev)
        return ops.gemm(a6, w6, bias, out, epilogue, resid=resid, **self._wrap(K), **gemm_kw)

    def embed_exact(self, images: torch.Tensor, shift: int = 0) -> torch.Tensor:
        pm = self.pm
        B = images.shape[0]
        P, D, T, R, seq = pm.num_patches, pm.dim, pm.n_cls, pm.n_reg, pm.seq_len + shift
        Kp = pm.w_patch.shape[1]
        t = self.terms
        w6 = pm.extra.get(f"w_patch_x{t}")
        if w6 is None:
            w6 = pm.extra[f"w_patch_x{t}"] = ops.split_weight(pm.extra["conv_weight"].reshape(D, -1), t)
        # the patch GEMM's row remap runs on the single-CTA kernel: the two-term operand is stored unwrapped, [lo | hi | hi]
        patches6 = (self.ws.get(f"patches_x{t}", (B * P, (6 if t == 3 else 3) * Kp), torch.bfloat16) if self.patches_ready
                    else self.patchify(images))
        x = self.ws.get("x", (B * seq, D), torch.float32)
        ops.gemm(patches6, w6, pm.b_patch, x, PK_EPI_BIAS_RESID_F32, resid=pm.pos,
                 rows_per_group=P, group_stride=seq, group_offset=T + R + shift, resid_is_pos=True, pos_offset=T + R)
        ops.fill_token_rows(x, B, seq, 0, pm.cls_tokens[:1], pm.pos)
        if T > 1:
            ops.fill_token_rows(x, B, seq, 1 + shift, pm.cls_tokens[1:], pm.pos, pos_offset=1)
        if R > 0:
            ops.fill_token_rows(x, B, seq, T + shift, pm.reg_tokens, pm.pos, pos_offset=T)
        return x

    def dense_block_exact(self, x: torch.Tensor, lw: LayerWeights, rows: int, batch: int, seq: int) -> None:
        """ViTBlock (vit.py:45-55) at fp32 accuracy, in place on the residual stream."""
        self.attn_part(x, lw, rows, batch, seq=seq)
        self.mlp_part(x, lw, rows)

    def _attn_part_exact(self, x, lw, rows, batch, seq, cu, max_len, rows_dev, rowscale, key_mult, extra_mult, expert=0,
                         out_scale=None, a6=None):
        pm, ws = self.pm, self.ws
        D = pm.dim
        dh = D // pm.heads
        w, aw = self._exact_weights(lw), lw.attn[expert]
        xr = x[:rows]
        o_scale = rowscale if out_scale is None else out_scale
        if (self.terms == 2 and cu is None and key_mult is None and extra_mult is None and dh == 64 and 17 <= seq <= 224):
            # bf16x2 fast path: the in-projection writes IEEE-half q | k | v, the tcgen05 attention takes them as they are and
            # writes its fp32 result already split [lo | hi] -- the out-projection's operand, with no pass in between
            if a6 is None:
                a6 = ws.get(f"a2_{D}", (rows, 2 * D), torch.bfloat16)
                ops.split(xr, a6, 2, SPLIT_LAYERNORM, lw.ln1_w, lw.ln1_b, lw.eps, rows=rows, rowscale=rowscale, rows_dev=rows_dev)
            qkv = ops.gemm(a6, w["w_qkv"][expert], aw.b_qkv, ws.get("qkv16", (rows, 3 * D), torch.float16), PK_EPI_BIAS_BF16,
                           m_dev=rows_dev, out_format=PK_OUT_F16, **self._wrap(D))
            att = ops.attention(qkv, ws.get("att_x2", (rows, 2 * D), torch.bfloat16), batch, pm.heads, dh, seq_len=seq, half_split=True)
            ops.gemm(att, w["w_o"][expert], aw.b_o, xr, PK_EPI_BIAS_RESID_F32, resid=xr, m_dev=rows_dev, rowscale=o_scale, **self._wrap(D))
            return
        qkv = self.gemm_exact(xr, rows, w["w_qkv"][expert], aw.b_qkv, ws.get("qkv32", (rows, 3 * D), torch.float32), PK_EPI_BIAS_F32,
                              SPLIT_LAYERNORM, lw.ln1_w, lw.ln1_b, lw.eps, in_scale=rowscale, rows_dev=rows_dev, m_dev=rows_dev, a6=a6)
        att = ops.attention_f32(qkv, ws.get("att32", (rows, D), torch.float32), batch, pm.heads, dh, seq,
                                cu_seqlens=cu, max_seq_len=max_len, key_mult=key_mult,
                                extra_kv=aw.b_qkv[D:] if extra_mult is not None else None, extra_mult=extra_mult)
        self.gemm_exact(att, rows, w["w_o"][expert], aw.b_o, xr, PK_EPI_BIAS_RESID_F32, resid=xr, rows_dev=rows_dev, m_dev=rows_dev,
                        rowscale=o_scale)

    def _mlp_part_exact(self, x, lw, rows, rows_dev, rowscale, residual: bool = True):
        ws = self.ws
        epi, res = (PK_EPI_BIAS_RESID_F32, x[:rows]) if residual else (PK_EPI_BIAS_F32, None)
        w, mw = self._exact_weights(lw), lw.mlp[0]
        F = mw.b_fc1.shape[0]
        xr = x[:rows]
        if self.terms == 2 and F % 64 == 0:
            # bf16x2 fast path: fc1's epilogue applies bias + GELU and writes the hidden row already split [lo | hi]
            D = xr.shape[-1]
            a2 = ws.get(f"a2_{D}", (rows, 2 * D), torch.bfloat16)
            ops.split(xr, a2, 2, SPLIT_LAYERNORM, lw.ln2_w, lw.ln2_b, lw.eps, rows=rows, rowscale=rowscale, rows_dev=rows_dev)
            hid = ops.gemm(a2, w["w_fc1"][0], mw.b_fc1, ws.get("hid_x2", (rows, 2 * F), torch.bfloat16), PK_EPI_BIAS_GELU_BF16,
                           m_dev=rows_dev, out_format=PK_OUT_BF16X2, **self._wrap(D))
            ops.gemm(hid, w["w_fc2"][0], mw.b_fc2, xr, epi, resid=res, m_dev=rows_dev, **self._wrap(F))
            return
        hid = self.gemm_exact(xr, rows, w["w_fc1"][0], mw.b_fc1, ws.get("hid32", (rows, F), torch.float32), PK_EPI_BIAS_F32,
                              SPLIT_LAYERNORM, lw.ln2_w, lw.ln2_b, lw.eps, in_scale=rowscale, rows_dev=rows_dev, m_dev=rows_dev)
        self.gemm_exact(hid, rows, w["w_fc2"][0], mw.b_fc2, xr, epi, SPLIT_GELU, resid=res, rows_dev=rows_dev, m_dev=rows_dev)

    # ---- LayerNorm fused across GEMMs (dense layers of ViT / RankViT)
    def fold_ok(self, rows: int) -> bool:
        """The fused chain runs on the CTA-pair GEMM: more than one 256-row tile and 16-byte aligned rows."""
        return LN_FOLD > 0 and rows > 256 and self.pm.dim % 8 == 0

    def fold_begin(self, x: torch.Tensor, rows: int):
        """bf16 copy + row statistics of a residual stream that did not come out of a producer GEMM."""
        D = self.pm.dim
        xb = self.ws.get("xb", (rows, D), torch.bfloat16)
        stats = self.ws.get("ln_stats", (rows, ops.gemm_row_stat_parts(D), 2), torch.float32)
        ops.row_stats_cast(x, xb, stats, rows)
        return xb, stats

    def dense_block_fused(self, x, lw: LayerWeights, rows: int, batch: int, seq: int, xb, stats, emit_last: bool) -> None:
        """ViTBlock (vit.py:45-55) without LayerNorm kernels: ln_1 / ln_2 are folded into the in-proj and fc1 GEMMs
        (gain in the weight, mean / rstd applied in the epilogue from per-row statistics) and the out-proj / fc2
        epilogues emit the bf16 copy and the statistics of the row they just produced."""
        pm, ws = self.pm, self.ws
        D = pm.dim
        aw, mw, f = lw.attn[0], lw.mlp[0], lw.extra["fold"]
        F = mw.w_fc1.shape[0]
        qkv = ops.gemm(xb, f["w_qkv"], f["c2_qkv"], ws.get("qkv", (rows, 3 * D), torch.bfloat16, zero=True), PK_EPI_BIAS_BF16,
                       ln_stats=stats, ln_c1=f["c1_qkv"], ln_dim=D, ln_eps=lw.eps)
        att = ops.attention(qkv, ws.get("att", (rows, D), torch.bfloat16), batch, pm.heads, D // pm.heads, seq_len=seq)
        if LN_FOLD >= 2:
            ops.gemm(att, aw.w_o, aw.b_o, x[:rows], PK_EPI_BIAS_RESID_F32, resid=x[:rows], xb_out=xb, row_stats=stats)
            hid = ops.gemm(xb, f["w_fc1"], f["c2_fc1"], ws.get("hid", (rows, F), torch.bfloat16), PK_EPI_BIAS_GELU_BF16,
                           ln_stats=stats, ln_c1=f["c1_fc1"], ln_dim=D, ln_eps=lw.eps)
        else:
            ops.gemm(att, aw.w_o, aw.b_o, x[:rows], PK_EPI_BIAS_RESID_F32, resid=x[:rows])
            a = ops.layernorm(x, lw.ln2_w, lw.ln2_b, lw.eps, ws.get("ln", (rows, D), torch.bfloat16), rows=rows)
            hid = ops.gemm(a, mw.w_fc1, mw.b_fc1, ws.get("hid", (rows, F), torch.bfloat16), PK_EPI_BIAS_GELU_BF16)
        if emit_last:
            ops.gemm(hid, mw.w_fc2, mw.b_fc2, x[:rows], PK_EPI_BIAS_RESID_F32, resid=x[:rows], xb_out=xb, row_stats=stats)
        else:
            ops.gemm(hid, mw.w_fc2, mw.b_fc2, x[:rows], PK_EPI_BIAS_RESID_F32, resid=x[:rows])

    def head(self, x: torch.Tensor, batch: int, seq: int, cu_seqlens=None, n_cls: Optional[int] = None) -> torch.Tensor:
        """Final LayerNorm on the class rows, class-token sum, linear head (vit.py:95,242-246), in fp32.  Batches of more than
        256 samples run the head as a tensor-core GEMM over 3-way split operands (the fp32-accurate form: 1e-6 of the fp32
        product), fed by a one-warp-per-sample LayerNorm; smaller ones use the fused CUDA-core kernel."""
        pm = self.pm
        n_cls = pm.n_cls if n_cls is None else n_cls
        logits = self.ws.get("logits", (batch, pm.num_classes), torch.float32)
        if HEAD_GEMM and batch > 256 and pm.num_classes % 8 == 0 and pm.dim % 8 == 0:
            w6 = pm.extra.get("head_w6")
            if w6 is None:
                w6 = pm.extra["head_w6"] = ops.split3_weight(pm.head_w)
            feat = ops.cls_features(x, batch, seq, n_cls, pm.ln_w, pm.ln_b, pm.ln_eps, self.ws.get("cls_feat", (batch, pm.dim), torch.float32),
                                    cu_seqlens=cu_seqlens)
            f6 = ops.split3(feat, self.ws.get("cls_feat6", (batch, 6 * pm.dim), torch.bfloat16))
            return ops.gemm(f6, w6, pm.head_b, logits, PK_EPI_BIAS_F32)
        return ops.cls_head(x, batch, seq, n_cls, pm.ln_w, pm.ln_b, pm.ln_eps, pm.head_w, pm.head_b, cu_seqlens=cu_seqlens, out=logits)

    def _const(self, name: str, builder):
        """Device constants of the ragged paths (initial cu_seqlens, token ids ...), built once per shape."""
        t = self.ws._bufs.get(("const", name))
        if t is None:
            t = builder()
            self.ws._bufs[("const", name)] = t
        return t

    # ---------------------------------------------------------------- plain ViT
    def vit(self, images: torch.Tensor) -> torch.Tensor:
        pm = self.pm
        B, seq = images.shape[0], pm.seq_len
        rows = B * seq
        if self.exact:
            x = self.embed_exact(images)
            for i, lw in enumerate(pm.layers):
                if lw.kind == "noise":
                    apply_noise(self.noise_draws[i], x, B, seq, self.sample_offset)
                else:
                    self.dense_block_exact(x, lw, rows, B, seq)
            return self.head(x, B, seq)
        fold_all = self.fold_ok(rows) and all("fold" in lw.extra for lw in pm.layers)
        x, fold = self.embed(images, emit_fold=fold_all)
        if fold_all:
            xb, stats = fold if fold is not None else self.fold_begin(x, rows)
            for i, lw in enumerate(pm.layers):
                self.dense_block_fused(x, lw, rows, B, seq, xb, stats, emit_last=i + 1 < len(pm.layers))
        else:
            for i, lw in enumerate(pm.layers):
                if lw.kind == "noise":
                    apply_noise(self.noise_draws[i], x, B, seq, self.sample_offset)
                else:
                    self.dense_block(x, lw, rows, B, seq)
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
        # (xb, stats) of the current residual stream when valid; the patch GEMM emits them for layer 0
        if self.exact:
            x, fold = self.embed_exact(images), None
        else:
            x, fold = self.embed(images, emit_fold=bool(pm.layers) and self.fold_ok(B * seq) and "fold" in pm.layers[0].extra)
        flip = 0
        L = len(pm.layers)
        for i, lw in enumerate(pm.layers):
            if lw.kind == "noise":
                apply_noise(self.noise_draws[i], x, B, seq, self.sample_offset)
                fold = None
                continue
            b = budgets.get(i, 1.0) if lw.kind == "rank" else 1.0
            if lw.kind == "rank" and b != 1 and seq > 1:            # nothing left to rank once only the class token survives
                n = seq - 1
                # rankvit.py:69-71: idx[:, :ceil(n*b)] -- a budget of 0 keeps the class token alone, a budget above 1 every token
                k = min(max(math.ceil(n * b), 0), n)
                scores = ops.token_norm_score(x, B, seq, ws.get(f"rank_scores_{n}", (B, n), torch.float32))
                kept = ws.get(f"rank_kept_{n}_{k}", (B, k), torch.int32)
                if k > 0:
                    ops.topk_select(scores, k, kept)
                y = ws.get(f"x_compact_{flip}", (B * (k + 1), D), torch.float32)
                flip ^= 1
                ops.gather_rows(x, kept, B, seq, y)
                if aux is not None:
                    aux.setdefault("scores", {})[i] = scores.clone()
                    aux.setdefault("kept", {})[i] = kept.clone()
                x, seq = y, k + 1
                fold = None                          # new rows: the bf16 copy / statistics are stale
            if aux is not None:
                aux.setdefault("seq_lens", []).append(seq)
            rows = B * seq
            if self.exact:
                self.dense_block_exact(x, lw, rows, B, seq)
            elif self.fold_ok(rows) and "fold" in lw.extra:
                if fold is None:
                    fold = self.fold_begin(x, rows)
                self.dense_block_fused(x, lw, rows, B, seq, fold[0], fold[1], emit_last=i + 1 < L)
            else:
                self.dense_block(x, lw, rows, B, seq)
                fold = None
        return self.head(x, B, seq)

    # ---------------------------------------------------------------- ResidualViT
    def residualvit(self, images: torch.Tensor, budget: float, aux: Optional[dict] = None) -> torch.Tensor:
        """Budget-token gating with real compaction (reference residualvit.py:587-616, block :197-260).
        Local row layout of a sample: [cls, budget token, live image rows ..., ghost slot]."""
        pm, ws = self.pm, self.ws
        ex = pm.extra
        if ex.get("dense_layout"):
            return self.residualvit_dense(images, budget, aux)
        abt = ex["add_budget_token"]
        B, D, dev = images.shape[0], pm.dim, images.device
        nb = 1 if abt else 0
        n_special = 1 + nb
        n_img = pm.seq_len - 1                                   # other class tokens / registers / patches are all gated
        seq0 = pm.seq_len + nb
        cap = seq0 + 1                                           # + ghost slot
        rows_cap = B * cap
        x = self.embed_exact(images, shift=nb) if self.exact else self.embed(images, shift=nb)
        if abt:
            if budget is None:
                raise AssertionError("Budget token not set. Call set_budget() before forward() to evaluate the model on a chosen budget.")
            if abt == "learnable":
                ops.fill_token_rows(x, B, seq0, 1, ex["budget_token_1"], None, scale=float(budget))       # residualvit.py:572-576
            elif abt == "learnable_interpolate":
                tok = ex["budget_token_1"] * float(budget) + ex["budget_token_2"] * (1.0 - float(budget))  # :577-580
                ops.fill_token_rows(x, B, seq0, 1, tok, None)
            else:
                ops.fill_token_rows(x, B, seq0, 1, None, None, scale=float(budget), n_tokens=1)           # :581-583
        cu = self._const(f"cu0_{B}_{seq0}", lambda: (torch.arange(B + 1, device=dev, dtype=torch.int32) * seq0))
        rows_dev = self._const(f"rows0_{B}_{seq0}", lambda: torch.tensor([B * seq0], device=dev, dtype=torch.int32))
        mult = self._const(f"ones_{rows_cap}", lambda: torch.ones(rows_cap, device=dev, dtype=torch.float32))
        tok_row0 = self._const(f"tokrow0_{B}_{seq0}_{n_special}", lambda: (
            torch.arange(B, device=dev, dtype=torch.int32)[:, None] * seq0 + n_special
            + torch.arange(n_img, device=dev, dtype=torch.int32)[None, :]).contiguous())
        tok_row = ws.get("res_tok_row", (B, n_img), torch.int32)
        tok_row.copy_(tok_row0)
        have_mult = False
        bufs = [ws.get("res_x0", (rows_cap, D), torch.float32), ws.get("res_x1", (rows_cap, D), torch.float32)]
        mults = [ws.get("res_mult0", (rows_cap,), torch.float32), ws.get("res_mult1", (rows_cap,), torch.float32)]
        cus = [ws.get("res_cu0", (B + 1,), torch.int32), ws.get("res_cu1", (B + 1,), torch.int32)]
        rdevs = [ws.get("res_rows0", (1,), torch.int32), ws.get("res_rows1", (1,), torch.int32)]
        mask = ws.get("res_mask", (rows_cap,), torch.float32)
        rowscale = ws.get("res_rowscale", (rows_cap,), torch.float32)
        dst_local = ws.get("res_dst", (rows_cap,), torch.int32)
        sample_of = ws.get("res_sample", (rows_cap,), torch.int32)
        new_len = ws.get("res_newlen", (B,), torch.int32)
        mdrop = ws.get("res_mdrop", (B,), torch.float32)
        thr_dev = ws.get("res_thr", (1,), torch.float32)
        ee = ex.get("ee_heads")
        if ee is not None and len(ee) != len(pm.layers):
            raise RuntimeError("EEResidualViT: one early-exit head per encoder layer is required (eeresidualvit.py:93-94)")
        exits = ws.get("ee_logits", (len(pm.layers) + 1, B, pm.num_classes), torch.float32) if ee is not None else None
        flip = 0
        for i, lw in enumerate(pm.layers):
            skip = lw.extra.get("skip")
            if skip == "attention+mlp":
                if not abt:
                    raise RuntimeError("ResidualViT 'attention+mlp' layers need a budget token (the reference raises a shape "
                                       "error here, residualvit.py:230-237)")
                g = lw.extra
                if g["gate_type"] == "gumbel" and abt != "learnable":
                    raise AssertionError("Gumbel gate does not support budget")
                thr_mode = 0 if abt == "learnable" else 1
                if thr_mode == 1:
                    ops.budget_mean_threshold(x, cu, B, 1, thr_dev)
                ops.residual_gate_plan(x, cu, mult, B, cap, n_special=n_special, budget_pos=1, gated=True, gate_w=g["gate_w"],
                                       gate_b=g["gate_b"], gate_temp=g["gate_temp"], gate_bias=g["gate_bias"],
                                       gate_type=0 if g["gate_type"] == "sigmoid" else 1, thr_mode=thr_mode, bt_w=g.get("bt_gate_w"),
                                       bt_b=g.get("bt_gate_b", 0.0), thr_dev=thr_dev,
                                       mask=mask, dst_local=dst_local, sample_of=sample_of, new_len=new_len, mdrop=mdrop)
                cu_out, rows_out, y, mult_out = cus[flip], rdevs[flip], bufs[flip], mults[flip]
                flip ^= 1
                ops.exclusive_scan(new_len, cu_out, rows_out)
                # block.mask (B, N_img, 1) is published by the same launch (the token -> row map moves to the compacted layout)
                mask_pub = torch.empty(B, n_img, 1, device=dev, dtype=torch.float32) if aux is not None else None
                ops.compact_rows(x, y, cu, cu_out, B, rows_cap, dst_local, sample_of, scale_in=mask, scale_out=rowscale,
                                 attrs=[(mult, mult_out)], ghost=True,
                                 publish=(tok_row, mask_pub, n_img) if aux is not None else None)
                if aux is not None:
                    aux.setdefault("masks", {})[i] = mask_pub
                    aux.setdefault("rows", {})[i] = rows_out.clone()
                x, cu, rows_dev, mult, have_mult = y, cu_out, rows_out, mult_out, True
                self.attn_part(x, lw, rows_cap, B, cu=cu, max_len=cap, rows_dev=rows_dev, rowscale=rowscale, key_mult=mult,
                               extra_mult=mdrop)
                self.mlp_part(x, lw, rows_cap, rows_dev=rows_dev, rowscale=rowscale)
                ops.residual_ghost(x, mult, cu, mdrop, g["mlp0"], B)
            elif skip in (None, "none"):
                self.attn_part(x, lw, rows_cap, B, cu=cu, max_len=cap, rows_dev=rows_dev, key_mult=mult if have_mult else None)
                self.mlp_part(x, lw, rows_cap, rows_dev=rows_dev)
            else:
                raise NotImplementedError(f"skip mode {skip!r}")
            if ee is not None:
                # early-exit head of this layer on the first class token = local row 0 of every sample (eeresidualvit.py:94)
                ln_w, ln_b, ln_eps, hw, hb = ee[i]
                ops.cls_head(x, B, 0, 1, ln_w, ln_b, ln_eps, hw, hb, cu_seqlens=cu, out=exits[i])
        if ee is not None:
            ops.cls_head(x, B, 0, 1, pm.ln_w, pm.ln_b, pm.ln_eps, pm.head_w, pm.head_b, cu_seqlens=cu, out=exits[len(pm.layers)])
            return exits
        return self.head(x, B, 0, cu_seqlens=cu, n_cls=1)

    def residualvit_dense(self, images: torch.Tensor, budget, aux: Optional[dict] = None) -> torch.Tensor:
        """ResidualViT on the dense masked row layout -- every token keeps its row, like in the reference -- for the
        configurations whose dropped rows do not stay identical: skip modes 'attention' (residualvit.py:130-157) and 'mlp'
        (:160-194), ``add_input`` (:239-242), two special tokens without a budget token, a NoiseBlock between blocks.
        Local row layout of a sample: [cls0, budget token (if any), other class tokens, registers, patches]."""
        pm, ws = self.pm, self.ws
        ex = pm.extra
        abt = ex["add_budget_token"]
        B, D, dev = images.shape[0], pm.dim, images.device
        nb = 1 if abt else 0
        # ResidualViTBlock.num_special_tokens (:99): the model builds its encoder without telling it about further class tokens
        # or registers (:452-467), so those are gated like image tokens and only the first class token is special
        n_special = next((int(lw.module.num_special_tokens) for lw in pm.layers if lw.kind == "residual"), 1)
        front = n_special + nb                                 # rows never gated: specials + the budget token
        seq = pm.seq_len + nb
        n_img = seq - front
        rows = B * seq
        x = self.embed_exact(images, shift=nb) if self.exact else self.embed(images, shift=nb)
        if abt:
            if budget is None:
                raise AssertionError("Budget token not set. Call set_budget() before forward() to evaluate the model on a chosen budget.")
            if abt == "learnable":
                ops.fill_token_rows(x, B, seq, 1, ex["budget_token_1"], None, scale=float(budget))
            elif abt == "learnable_interpolate":
                ops.fill_token_rows(x, B, seq, 1, ex["budget_token_1"] * float(budget) + ex["budget_token_2"] * (1.0 - float(budget)), None)
            else:
                ops.fill_token_rows(x, B, seq, 1, None, None, scale=float(budget), n_tokens=1)
        cu = self._const(f"cu0_{B}_{seq}", lambda: (torch.arange(B + 1, device=dev, dtype=torch.int32) * seq))
        ones = self._const(f"ones_{rows}", lambda: torch.ones(rows, device=dev, dtype=torch.float32))
        t = ws.get("resd_x1", (rows, D), torch.float32)
        mask = ws.get("resd_mask", (rows,), torch.float32)
        scratch_i = [ws.get(f"resd_i{j}", (rows,), torch.int32) for j in range(2)]
        new_len, mdrop = ws.get("resd_newlen", (B,), torch.int32), ws.get("resd_mdrop", (B,), torch.float32)
        thr_dev = ws.get("resd_thr", (1,), torch.float32)
        ee = ex.get("ee_heads")
        if ee is not None and len(ee) != len([lw for lw in pm.layers if lw.kind != "noise"]):
            raise RuntimeError("EEResidualViT: one early-exit head per encoder layer is required (eeresidualvit.py:93-94)")
        n_blocks = len(ee) if ee is not None else 0
        exits = ws.get("ee_logits", (n_blocks + 1, B, pm.num_classes), torch.float32) if ee is not None else None

        def gate(src, g, mode: str) -> None:
            """block.mask of this layer for the rows of ``src`` (ResidualGate.forward, residualvit.py:47-74) -> ``mask`` (1 on
            the special / budget rows)."""
            sig = g["gate_type"] == "sigmoid"
            if abt and mode == "attention+mlp" and abt == "learnable":
                thr_mode = 0                                   # threshold = sigmoid(budget_token_gate(budget token)) (:212)
            elif abt:
                if not sig:
                    raise AssertionError("Gumbel gate does not support budget")
                thr_mode = 1                                   # budget = mean of the budget token over the batch (:142,177,208)
                ops.budget_mean_threshold(src, cu, B, 1, thr_dev)
            else:
                thr_mode = 2                                   # the gate's own threshold (:69)
            ops.residual_gate_plan(src, cu, ones, B, seq, n_special=front, budget_pos=1 if abt else -1, gated=True,
                                   gate_w=g["gate_w"], gate_b=g["gate_b"], gate_temp=g["gate_temp"], gate_bias=g["gate_bias"],
                                   gate_type=0 if sig else 1, thr_mode=thr_mode, bt_w=g.get("bt_gate_w"), bt_b=g.get("bt_gate_b", 0.0),
                                   thr_dev=thr_dev, thr_const=float(g["gate_threshold"]), mask=mask, dst_local=scratch_i[0],
                                   sample_of=scratch_i[1], new_len=new_len, mdrop=mdrop)

        def publish(i: int) -> None:
            if aux is not None:
                aux.setdefault("masks", {})[i] = mask.view(B, seq)[:, front:].reshape(B, n_img, 1).clone()

        blk_i = 0
        for i, lw in enumerate(pm.layers):
            if lw.kind == "noise":
                apply_noise(self.noise_draws[i], x, B, seq, self.sample_offset)
                continue
            skip, g = lw.extra.get("skip"), lw.extra
            if skip == "attention+mlp":
                if front != 2:
                    raise RuntimeError("ResidualViT 'attention+mlp' layers need exactly two ungated tokens, i.e. a budget token "
                                       "(the reference's forward mask hard-codes them, residualvit.py:230-235)")
                if g["add_input"] and abt:
                    raise RuntimeError("add_input with a budget token: the reference adds a tensor one token short here "
                                       "(residualvit.py:239-242)")
                gate(x, g, skip)
                publish(i)
                ops.row_scale_add(t, x, mask, rows)                                     # masked_input (:220-227)
                self.attn_part(t, lw, rows, B, seq=seq, rowscale=mask)
                self.mlp_part(t, lw, rows, rowscale=mask)
                if g["add_input"]:
                    ops.row_scale_add(t, x, mask, rows, one_minus=True, accumulate=True)   # + img_tokens * (1 - mask)
                x, t = t, x
            elif skip == "attention":
                if abt:
                    raise RuntimeError("ResidualViT skip mode 'attention' cannot run with a budget token (the reference adds "
                                       "tensors of different lengths, residualvit.py:137-152)")
                gate(x, g, skip)
                publish(i)
                ops.row_scale_add(t, x, mask, rows)                                     # masked_input (:145-148)
                self.attn_part(t, lw, rows, B, seq=seq)                                 # t = masked_input + attention(ln_1(masked_input))
                ops.row_scale_add(t, x, mask, rows, one_minus=True, accumulate=True)    # ... + input instead (:150-153)
                self.mlp_part(t, lw, rows, residual=False)                              # returns mlp(ln_2(x)) alone (:155-157)
                x, t = t, x
            elif skip == "mlp":
                if g["add_input"] and abt:
                    raise RuntimeError("add_input with a budget token: the reference adds a tensor one token short here "
                                       "(residualvit.py:189-192)")
                self.attn_part(x, lw, rows, B, seq=seq)                                 # :162-165
                gate(x, g, skip)
                publish(i)
                ops.row_scale_add(t, x, mask, rows)                                     # masked_input (:178-184)
                self.mlp_part(t, lw, rows, residual=False)                              # :186-187
                if g["add_input"]:
                    ops.row_scale_add(t, x, mask, rows, one_minus=True, accumulate=True)
                x, t = t, x
            elif skip in (None, "none"):
                self.attn_part(x, lw, rows, B, seq=seq)
                self.mlp_part(x, lw, rows)
            else:
                raise NotImplementedError(f"skip mode {skip!r}")
            if ee is not None:
                ln_w, ln_b, ln_eps, hw, hb = ee[blk_i]
                ops.cls_head(x, B, seq, 1, ln_w, ln_b, ln_eps, hw, hb, out=exits[blk_i])
            blk_i += 1
        if ee is not None:
            ops.cls_head(x, B, seq, 1, pm.ln_w, pm.ln_b, pm.ln_eps, pm.head_w, pm.head_b, out=exits[n_blocks])
            return exits
        if pm.n_cls > 1 and nb:
            # the class tokens sit at local rows 0, 2, 3, ...: gather them behind each other for the sum readout (:609-611)
            pick = self._const(f"cls_pick_{B}_{pm.n_cls}", lambda: torch.arange(1, pm.n_cls, device=dev, dtype=torch.int32).repeat(B, 1).contiguous())
            y = ops.gather_rows(x, pick, B, seq, ws.get("resd_cls", (B * pm.n_cls, D), torch.float32))
            return self.head(y, B, pm.n_cls, n_cls=pm.n_cls)
        return self.head(x, B, seq, n_cls=pm.n_cls)

    # ---------------------------------------------------------------- AdaViT (A-ViT)
    def adavit(self, images: torch.Tensor, aux: Optional[dict] = None, early_exit: bool = True) -> torch.Tensor:
        """ACT halting with real token removal (reference adavit.py:140-219): halted tokens leave the packed
        batch and survive only as a virtual bias key; a sample retires once its class token halts."""
        pm, ws = self.pm, self.ws
        ex = pm.extra
        B, D, seq, dev = images.shape[0], pm.dim, pm.seq_len, images.device
        rows_cap = B * seq
        x = self.embed_exact(images) if self.exact else self.embed(images)
        cu = self._const(f"cu0_{B}_{seq}", lambda: (torch.arange(B + 1, device=dev, dtype=torch.int32) * seq))
        rows_dev = self._const(f"rows0_{B}_{seq}", lambda: torch.tensor([B * seq], device=dev, dtype=torch.int32))
        tokid0 = self._const(f"tokid0_{B}_{seq}", lambda: torch.arange(seq, device=dev, dtype=torch.float32).repeat(B).contiguous())
        xs = [x, ws.get("avit_x1", (rows_cap, D), torch.float32)]
        cs = [ws.get("avit_c0", (rows_cap,), torch.float32), ws.get("avit_c1", (rows_cap,), torch.float32)]
        Rs = [ws.get("avit_R0", (rows_cap,), torch.float32), ws.get("avit_R1", (rows_cap,), torch.float32)]
        toks = [ws.get("avit_tok0", (rows_cap,), torch.float32), ws.get("avit_tok1", (rows_cap,), torch.float32)]
        cus = [ws.get("avit_cu0", (B + 1,), torch.int32), ws.get("avit_cu1", (B + 1,), torch.int32)]
        rdevs = [ws.get("avit_rows0", (1,), torch.int32), ws.get("avit_rows1", (1,), torch.int32)]
        cs[0].zero_()
        Rs[0].fill_(1.0)
        toks[0].copy_(tokid0)
        out_acc = ws.get("avit_out", (B, D), torch.float32)
        out_acc.zero_()
        rho = torch.zeros(B, seq, device=dev, dtype=torch.float32)
        counter = torch.ones(B, seq, device=dev, dtype=torch.float32)
        dst_local = ws.get("avit_dst", (rows_cap,), torch.int32)
        sample_of = ws.get("avit_sample", (rows_cap,), torch.int32)
        new_len = ws.get("avit_newlen", (B,), torch.int32)
        n_halted = ws.get("avit_nhalted", (B,), torch.float32)
        cur = 0
        L = len(pm.layers)
        for i, lw in enumerate(pm.layers):
            if i == 0:
                # nothing has halted yet: every sample still has all its tokens, so the first block's attention is the dense
                # uniform-sequence kernel (135 us per 512 x 12 heads x 197 tokens; the ragged kernels 245 / 395 us)
                self.attn_part(xs[cur], lw, rows_cap, B, seq=seq)
            else:
                self.attn_part(xs[cur], lw, rows_cap, B, cu=cu, max_len=seq, rows_dev=rows_dev, extra_mult=n_halted)
            self.mlp_part(xs[cur], lw, rows_cap, rows_dev=rows_dev)
            if aux is not None:
                aux.setdefault("rows", []).append(rows_dev.clone())
            # the halting gate's scale / centre are plain attributes of the live block (adavit.py:32-33): read per forward
            ops.avit_halt_plan(xs[cur], cu, B, seq, cs[cur], Rs[cur], toks[cur], gate_scale=float(lw.module.gate_scale),
                               gate_center=float(lw.module.gate_center), eps=ex["eps"], last_layer=(i == L - 1), early_exit=early_exit,
                               out_acc=out_acc, rho=rho, counter=counter, dst_local=dst_local, sample_of=sample_of,
                               new_len=new_len, n_halted=n_halted)
            if i == L - 1:
                break
            nxt = cur ^ 1
            ops.exclusive_scan(new_len, cus[nxt], rdevs[nxt])
            ops.compact_rows(xs[cur], xs[nxt], cu, cus[nxt], B, rows_cap, dst_local, sample_of,
                             attrs=[(cs[cur], cs[nxt]), (Rs[cur], Rs[nxt]), (toks[cur], toks[nxt])])
            cu, rows_dev, cur = cus[nxt], rdevs[nxt], nxt
        if aux is not None:
            aux["rho_token"], aux["counter_token"] = rho, counter
        return self.head(out_acc, B, 1, n_cls=1)

    # ---------------------------------------------------------------- MoE ViT
    def moevit(self, images: torch.Tensor, aux: Optional[dict] = None) -> torch.Tensor:
        """Expert MLPs computed only for the tokens routed to them (reference moevit.py:49-61 evaluates every
        expert on every token and selects with a one-hot einsum)."""
        pm, ws = self.pm, self.ws
        B, seq, D, dev = images.shape[0], pm.seq_len, pm.dim, images.device
        rows = B * seq
        x = self.embed_exact(images) if self.exact else self.embed(images)
        for i, lw in enumerate(pm.layers):
            if lw.kind == "noise":
                apply_noise(self.noise_draws[i], x, B, seq, self.sample_offset)
                continue
            EA = len(lw.attn)
            if EA == 1:
                self.attn_part(x, lw, rows, B, seq=seq)
            else:
                # AttentionMoE (moevit.py:71-102): every expert attends over ALL tokens of the sample (its own K / V), and a
                # token takes the output of its arg-max expert.  Each expert's in-proj + attention runs densely, and the
                # out-proj residual epilogue adds only the rows routed to it (rowscale = one-hot gate), which is the
                # reference's stack + one-hot einsum without materialising (E, B, N, D).
                expert = ws.get("amoe_expert", (rows,), torch.int32)
                ops.moe_route(x, lw.ln1_w, lw.ln1_b, lw.eps, lw.extra["attn_gate_w"], lw.extra["attn_gate_b"], rows, expert,
                              ws.get(f"moe_off_{EA}", (EA + 1,), torch.int32), ws.get(f"moe_cnt_{EA}", (EA,), torch.int32),
                              ws.get("moe_src", (rows,), torch.int32),
                              scratch=ws.get("moe_sort_scratch", (ops.MOE_SORT_SCRATCH_INTS,), torch.int32))
                onehot = ops.expert_onehot(expert, EA, ws.get(f"amoe_onehot_{EA}", (EA, rows), torch.float32), rows)
                if self.exact:
                    # every expert reads the same split LN1 rows; they are produced by the first expert's in-proj call
                    sw = ops.split_width(self.terms)
                    a6 = ws.get(f"a{self.terms}_shared_{D}", (rows, sw * D), torch.bfloat16)     # not the buffer the out-proj splits into
                    ops.split(x, a6, self.terms, SPLIT_LAYERNORM, lw.ln1_w, lw.ln1_b, lw.eps, rows=rows)
                    for e in range(EA):
                        self._attn_part_exact(x, lw, rows, B, seq, None, 0, None, None, None, None, expert=e,
                                              out_scale=onehot[e], a6=a6)
                a = None if self.exact else ops.layernorm(x, lw.ln1_w, lw.ln1_b, lw.eps, ws.get("ln", (rows, D), torch.bfloat16), rows=rows)
                for e, aw in enumerate(() if self.exact else lw.attn):
                    qkv = ops.gemm(a, aw.w_qkv, aw.b_qkv, ws.get("qkv", (rows, 3 * D), torch.bfloat16, zero=True), PK_EPI_BIAS_BF16)
                    att = ops.attention(qkv, ws.get("att", (rows, D), torch.bfloat16), B, pm.heads, D // pm.heads, seq_len=seq)
                    ops.gemm(att, aw.w_o, aw.b_o, x, PK_EPI_BIAS_RESID_F32, resid=x, rowscale=onehot[e])
                if aux is not None:
                    aux.setdefault("attn_expert", {})[i] = expert.view(B, seq).clone()
            E = len(lw.mlp)
            if E == 1:
                self.mlp_part(x, lw, rows)
                continue
            expert = ws.get("moe_expert", (rows,), torch.int32)
            offsets = ws.get(f"moe_off_{E}", (E + 1,), torch.int32)
            counts = ws.get(f"moe_cnt_{E}", (E,), torch.int32)
            src_of = ws.get("moe_src", (rows,), torch.int32)
            ops.moe_route(x, lw.ln2_w, lw.ln2_b, lw.eps, lw.extra["mlp_gate_w"], lw.extra["mlp_gate_b"], rows, expert, offsets,
                          counts, src_of, scratch=ws.get("moe_sort_scratch", (ops.MOE_SORT_SCRATCH_INTS,), torch.int32))
            F = lw.mlp[0].w_fc1.shape[0]
            y_sorted = ws.get("moe_y", (rows, D), torch.float32)
            if self.exact:
                # same expert-sorted segments at fp32 accuracy: split(LN2) gathered into sorted order, per-expert fc1 over its
                # segment, one exact-GELU split of all hidden rows, per-expert fc2
                w6 = self._exact_weights(lw)
                t, sw = self.terms, ops.split_width(self.terms)
                a6 = ws.get(f"a{t}_shared_{D}", (rows, sw * D), torch.bfloat16)
                ops.split(x, a6, t, SPLIT_LAYERNORM, lw.ln2_w, lw.ln2_b, lw.eps, rows=rows, row_index=src_of)
                if t == 2 and F % 64 == 0:
                    # bf16x2: every expert's fc1 epilogue writes its segment of the hidden rows already split [lo | hi]
                    h6 = ws.get("hid_x2", (rows, 2 * F), torch.bfloat16)
                    for e, mw in enumerate(lw.mlp):
                        ops.gemm(a6, w6["w_fc1"][e], mw.b_fc1, h6, PK_EPI_BIAS_GELU_BF16, m_dev=counts[e:e + 1],
                                 row_begin_dev=offsets[e:e + 1], out_format=PK_OUT_BF16X2, **self._wrap(D))
                else:
                    hid32 = ws.get("hid32", (rows, F), torch.float32)
                    for e, mw in enumerate(lw.mlp):
                        ops.gemm(a6, w6["w_fc1"][e], mw.b_fc1, hid32, PK_EPI_BIAS_F32, m_dev=counts[e:e + 1], row_begin_dev=offsets[e:e + 1],
                                 **self._wrap(D))
                    h6 = ops.split(hid32, ws.get(f"a{t}_{F}", (rows, sw * F), torch.bfloat16), t, SPLIT_GELU, rows=rows)
                for e, mw in enumerate(lw.mlp):
                    ops.gemm(h6, w6["w_fc2"][e], mw.b_fc2, y_sorted, PK_EPI_BIAS_F32, m_dev=counts[e:e + 1], row_begin_dev=offsets[e:e + 1],
                             **self._wrap(F))
                ops.scatter_add_rows(x, y_sorted, src_of, rows)
                if aux is not None:
                    aux.setdefault("mlp_expert", {})[i] = expert.view(B, seq).clone()
                continue
            a = ops.layernorm(x, lw.ln2_w, lw.ln2_b, lw.eps, ws.get("ln", (rows, D), torch.bfloat16), rows=rows, row_index=src_of)
            hid = ws.get("hid", (rows, F), torch.bfloat16)
            # Expert-sorted rows: each expert's segment [offsets[e], offsets[e] + counts[e]) runs through the CTA-pair GEMMs
            # (device-side segment start and length); the fc2 outputs land in sorted order and one gather-add un-permutes
            # them into the residual stream.
            fused = MOE_FUSED_SCATTER and rows > 256 and D % 8 == 0
            if fused and MOE_GROUPED:
                # ONE launch per projection over all expert segments: the tile scheduler of the CTA-pair GEMM walks the
                # segments [offsets[e], offsets[e + 1]) and takes expert e's rows of the stacked weight / bias
                st = lw.extra.get("moe_stacked")
                if st is None:
                    st = lw.extra["moe_stacked"] = (torch.cat([m.w_fc1 for m in lw.mlp]).contiguous(), torch.cat([m.b_fc1 for m in lw.mlp]).contiguous(),
                                                    torch.cat([m.w_fc2 for m in lw.mlp]).contiguous(), torch.cat([m.b_fc2 for m in lw.mlp]).contiguous())
                ops.gemm(a, st[0], st[1], hid, PK_EPI_BIAS_GELU_BF16, group_offsets=offsets, n_groups=E, cta_pair=2)
                ops.gemm(hid, st[2], st[3], x, PK_EPI_BIAS_RESID_F32, resid=x, out_row_index=src_of, group_offsets=offsets, n_groups=E,
                         cta_pair=2)
                if aux is not None:
                    aux.setdefault("mlp_expert", {})[i] = expert.view(B, seq).clone()
                continue
            for e, mw in enumerate(lw.mlp):
                ops.gemm(a, mw.w_fc1, mw.b_fc1, hid, PK_EPI_BIAS_GELU_BF16, m_dev=counts[e:e + 1], row_begin_dev=offsets[e:e + 1])
                if fused:
                    # the fc2 epilogue un-permutes and adds into the residual stream itself: x[src_of[r]] += fc2(hid[r]) + b
                    # (vector reductions at L2; src_of is a permutation) -- no sorted fp32 output, no gather-add pass
                    ops.gemm(hid, mw.w_fc2, mw.b_fc2, x, PK_EPI_BIAS_RESID_F32, resid=x, m_dev=counts[e:e + 1],
                             row_begin_dev=offsets[e:e + 1], out_row_index=src_of, cta_pair=2)
                else:
                    ops.gemm(hid, mw.w_fc2, mw.b_fc2, y_sorted, PK_EPI_BIAS_F32, m_dev=counts[e:e + 1], row_begin_dev=offsets[e:e + 1])
            if not fused:
                ops.scatter_add_rows(x, y_sorted, src_of, rows)
            if aux is not None:
                aux.setdefault("mlp_expert", {})[i] = expert.view(B, seq).clone()
        return self.head(x, B, seq, n_cls=1)
