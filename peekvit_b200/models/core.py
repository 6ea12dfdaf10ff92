"""``nn.Module`` mirrors of the reference model classes (the drop-in boundary).

Same class names, constructor kwargs, parameter names/shapes (so reference checkpoints and
``load_state`` round-trip unchanged), public attributes (``set_budget``, ``current_budget``,
``encoder.layers`` as a mutable ``nn.Sequential``, per-block ``.mask`` …) — but ``forward``
runs the hand-written sm_100a kernels through the C ABI instead of ATen ops, and really skips
the tokens a budget drops.  The submodules (``nn.LayerNorm``, ``nn.MultiheadAttention``,
``nn.Linear``, ``nn.Conv2d``) are used as *parameter containers only*: their own ``forward``
is never called.  Inference only: there is no backward on this path and no CPU fallback.

Reference classes mirrored: ``VisionTransformer`` (models/vit.py:100-315),
``RankVisionTransformer`` (models/rankvit.py:156-339), ``ResidualVisionTransformer``
(models/residualvit.py:352-694), ``AdaptiveVisionTransformer`` (models/adavit.py:225-433),
``VisionTransformerMoE`` (models/moevit.py:191-315) and their blocks/encoders.
"""
from __future__ import annotations

import math
from abc import ABC
from typing import List, Literal, Optional, Union

import random

import torch
from torch import nn

from .. import runner


def _no_forward(self, *a, **k):
    raise RuntimeError(
        f"{type(self).__name__} is a parameter container on the B200 path; the block math runs fused inside the "
        "model's forward (peekvit_b200.engine). Call the top-level model instead.")


# ------------------------------------------------------------------------------ building blocks
class MLP(nn.Module):
    """Parameters of the reference MLP (models/blocks.py:74-84): fc1 -> GELU(erf) -> fc2."""

    def __init__(self, hidden_dim, mlp_dim):
        super().__init__()
        self.fc1 = nn.Linear(hidden_dim, mlp_dim)
        self.fc2 = nn.Linear(mlp_dim, hidden_dim)

    forward = _no_forward


class SelfAttention(nn.Module):
    """Parameters of the reference SelfAttention (models/blocks.py:88-95): a packed-QKV
    ``nn.MultiheadAttention(batch_first=True)`` under the attribute ``self_attention``."""

    def __init__(self, input_dim, num_heads, dropout=0.0):
        super().__init__()
        self.self_attention = nn.MultiheadAttention(input_dim, num_heads, batch_first=True, dropout=dropout)

    forward = _no_forward


class NoiseBlock(nn.Module):
    """Drop-in for reference models/blocks.py:100-188: Gaussian noise at a signal-to-noise ratio (dB) or token dropping,
    spliced into ``encoder.layers`` by ``add_noise`` (utils/utils.py:162-191).  Same constructor, setters and attributes;
    inside a model forward the block runs as one kernel on the residual stream (peekvit_b200.engine), and calling it directly
    on a (B, N, D) CUDA tensor does the same out of place.  Noise is drawn with ``torch.randn`` / ``torch.randperm`` from
    torch's generators exactly like the reference, so seeding behaves the same."""

    def __init__(self, noise_type: Literal["gaussian", "token_drop"] = "gaussian", snr=None, std=None, prob=None):
        super().__init__()
        self.noise_type = noise_type
        self.snr_db, self.std, self.prob = snr, std, prob
        if not any([snr, std, prob]):
            print("Lazy initialization of noise block. Please set the noise parameters using set_snr, set_std or set_prob "
                  "before using the block.")
        if std:
            raise ValueError("std is not supported anymore. Please use snr instead.")

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        from .. import engine
        if not x.is_cuda:
            raise RuntimeError("peekvit_b200 has no CPU path: NoiseBlock needs a CUDA tensor")
        torch._assert(x.dim() == 3, f"Expected (batch_size, seq_length, hidden_dim) got {x.shape}")
        B, N, D = x.shape
        y = x.detach().to(torch.float32).contiguous().clone().view(B * N, D)
        engine.apply_noise(engine.draw_block_noise(self, B, N, D, x.device), y, B, N)
        return y.view(B, N, D)

    def set_snr(self, snr: float):
        assert self.noise_type == "gaussian"
        self.snr_db, self.std, self.prob = snr, None, None

    def set_prob(self, prob: float):
        assert self.noise_type == "token_drop"
        self.snr_db, self.std, self.prob = None, None, prob

    def set_value(self, value: float):
        if self.noise_type == "gaussian":
            self.set_snr(value)
        else:
            self.set_prob(value)


class _BlockBase(nn.Module):
    def __init__(self, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, ln_eps=1e-5):
        super().__init__()
        self.num_heads, self.hidden_dim, self.mlp_dim = num_heads, hidden_dim, mlp_dim
        self.ln_1 = nn.LayerNorm(hidden_dim, eps=ln_eps)
        self.self_attention = SelfAttention(hidden_dim, num_heads, attention_dropout)
        self.dropout = nn.Dropout(dropout)
        self.ln_2 = nn.LayerNorm(hidden_dim, eps=ln_eps)
        self.mlp = MLP(hidden_dim=hidden_dim, mlp_dim=mlp_dim)

    forward = _no_forward


class ViTBlock(_BlockBase):
    """models/vit.py:19-55."""


class RankViTBlock(_BlockBase):
    """models/rankvit.py:22-101: carries ``current_budget`` (1.0 = keep everything)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.sort = False
        self.current_budget = 1.0

    def set_budget(self, budget: float):
        self.current_budget = budget


class ResidualModule(ABC, nn.Module):
    """Marker base class looked up by reference utils.get_forward_masks (utils/utils.py:100-122)."""


class ResidualGate(nn.Module):
    """Parameters of models/residualvit.py:21-74 (Linear(D,1) + gate settings)."""

    def __init__(self, hidden_dim, threshold: Union[float, str] = 0.5, temp=1.0, gate_type="gumbel", sigmoid_bias: float = 10.0):
        super().__init__()
        self.projection = nn.Linear(hidden_dim, 1)
        self.temp, self.gate_type, self.sigmoid_bias = temp, gate_type, sigmoid_bias
        if gate_type not in ("gumbel", "sigmoid"):
            raise ValueError(f"Unknown gate type {gate_type}")
        if gate_type == "gumbel" and threshold != 0.5:
            raise ValueError("Gumbel gate cannot have a threshold different from 0.5")
        if isinstance(threshold, float):
            self.threshold = threshold
        elif threshold == "learnable":
            self.threshold = nn.Parameter(torch.tensor(0.5))

    forward = _no_forward


class ResidualViTBlock(ResidualModule):
    """models/residualvit.py:81-273 (LayerNorm eps 1e-6).  After every forward ``.mask`` holds the
    soft gate values ``(B, N_img, 1)`` exactly as the reference publishes them."""

    def __init__(self, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, temp: float = 1.0, add_input: bool = False,
                 num_class_tokens: int = 1, num_registers: int = 0,
                 skip: Literal["attention", "mlp", "attention+mlp", "none"] = None,
                 gate_type: Literal["gumbel", "sigmoid"] = "gumbel", gate_bias: float = 10.0, gate_threshold: float = 0.5,
                 budget_token: Union[bool, List, Literal["learnable"]] = False):
        super().__init__()
        self.num_heads, self.hidden_dim, self.mlp_dim = num_heads, hidden_dim, mlp_dim
        self.budget_token = budget_token
        self.num_special_tokens = num_class_tokens + num_registers
        self.gate_type, self.skip = gate_type, skip
        self.mask = None
        if skip in {"attention", "mlp", "attention+mlp"}:
            self.temp, self.add_input = temp, add_input
            self.residual_gate = ResidualGate(hidden_dim, threshold=gate_threshold, temp=temp, gate_type=gate_type,
                                              sigmoid_bias=gate_bias)
        else:
            self.add_input = False
        self.ln_1 = nn.LayerNorm(hidden_dim, eps=1e-06)
        self.self_attention = SelfAttention(hidden_dim, num_heads, dropout=attention_dropout)
        self.dropout = nn.Dropout(dropout)
        self.ln_2 = nn.LayerNorm(hidden_dim, eps=1e-06)
        self.mlp = MLP(hidden_dim=hidden_dim, mlp_dim=mlp_dim)
        if self.budget_token == "learnable":
            self.budget_token_gate = nn.Linear(hidden_dim, 1)

    forward = _no_forward


class AViTBlock(_BlockBase):
    """models/adavit.py:21-80: halting score sigmoid(x[...,0]*gate_scale - gate_center)."""

    def __init__(self, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, gate_scale: float = 10, gate_center: float = 30):
        super().__init__(num_heads, hidden_dim, mlp_dim, dropout, attention_dropout)
        self.gate_scale, self.gate_center = gate_scale, gate_center


class MoE(ABC, nn.Module):
    """Marker base class looked up by reference utils.get_last_forward_gates (utils/utils.py:76-94)."""


class TopKGate(nn.Module):
    """models/moevit.py:23-32: Linear(D,E); eval routing = one-hot arg-max."""

    def __init__(self, input_dim, num_experts):
        super().__init__()
        self.gate = nn.Linear(input_dim, num_experts)

    forward = _no_forward


class MLPMoE(MoE):
    """models/moevit.py:36-67."""

    def __init__(self, hidden_dim, mlp_dim, num_experts):
        super().__init__()
        self.gating_network = TopKGate(hidden_dim, num_experts)
        self.num_experts = num_experts
        self.experts = nn.ModuleList([MLP(hidden_dim, mlp_dim) for _ in range(num_experts)])
        self.gating_probs = None

    forward = _no_forward


class AttentionMoE(MoE):
    """models/moevit.py:69-102."""

    def __init__(self, input_dim, num_heads, num_experts, dropout=0.0):
        super().__init__()
        self.gating_network = TopKGate(input_dim, num_experts)
        self.num_experts = num_experts
        self.experts = nn.ModuleList([SelfAttention(input_dim, num_heads=num_heads, dropout=dropout) for _ in range(num_experts)])
        self.gating_probs = None

    forward = _no_forward


class ViTBlockMoE(nn.Module):
    """models/moevit.py:105-141."""

    def __init__(self, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, mlp_num_experts: int = 1, attn_num_experts: int = 1):
        super().__init__()
        self.num_heads = num_heads
        self.ln_1 = nn.LayerNorm(hidden_dim)
        self.self_attention = AttentionMoE(hidden_dim, num_heads, attn_num_experts, attention_dropout)
        self.dropout = nn.Dropout(dropout)
        self.ln_2 = nn.LayerNorm(hidden_dim)
        self.mlp = MLPMoE(hidden_dim=hidden_dim, mlp_dim=mlp_dim, num_experts=mlp_num_experts)

    forward = _no_forward


# ------------------------------------------------------------------------------ encoders
class _EncoderBase(nn.Module):
    def __init__(self, seq_length, hidden_dim, dropout, blocks, module_list=False):
        super().__init__()
        self.pos_embedding = nn.Parameter(torch.empty(1, seq_length, hidden_dim).normal_(std=0.02))
        self.dropout = nn.Dropout(dropout)
        self.layers = nn.ModuleList(blocks) if module_list else nn.Sequential(*blocks)
        self.ln = nn.LayerNorm(hidden_dim)

    forward = _no_forward


class ViTEncoder(_EncoderBase):
    """models/vit.py:59-95."""

    def __init__(self, seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout):
        super().__init__(seq_length, hidden_dim, dropout,
                         [ViTBlock(num_heads, hidden_dim, mlp_dim, dropout, attention_dropout) for _ in range(num_layers)])


class RankViTEncoder(_EncoderBase):
    """models/rankvit.py:105-149: RankViTBlock at the indices in ``rankvit_layers``."""

    def __init__(self, seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, rankvit_layers=None):
        blocks = [(RankViTBlock if i in rankvit_layers else ViTBlock)(num_heads, hidden_dim, mlp_dim, dropout, attention_dropout)
                  for i in range(num_layers)]      # TypeError when rankvit_layers is None, like the reference (rankvit.py:126)
        super().__init__(seq_length, hidden_dim, dropout, blocks)


class ResidualViTEncoder(_EncoderBase):
    """models/residualvit.py:278-348."""

    def __init__(self, seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout,
                 residual_layers: Optional[List] = None, add_input: bool = False, num_class_tokens: int = 1, num_registers: int = 0,
                 gate_type="gumbel", gate_temp: float = 1.0, gate_bias: float = 10.0, gate_threshold: float = 0.5,
                 budget_token: Union[bool, List, Literal["learnable"]] = False):
        blocks = [ResidualViTBlock(num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, skip=residual_layers[i],
                                   add_input=add_input, num_class_tokens=num_class_tokens, num_registers=num_registers,
                                   gate_type=gate_type, temp=gate_temp, gate_bias=gate_bias, gate_threshold=gate_threshold,
                                   budget_token=budget_token) for i in range(num_layers)]
        super().__init__(seq_length, hidden_dim, dropout, blocks)
        self.num_layers, self.num_class_tokens, self.num_registers = num_layers, num_class_tokens, num_registers
        self.num_special_tokens = num_class_tokens + num_registers
        self.budget_token = budget_token
        self.num_budget_tokens = 0 if not budget_token else 1


class EEResidualViTEncoder(ResidualViTEncoder):
    """models/eeresidualvit.py:17-96: the ResidualViT encoder plus ``early_exit_heads``, one
    ``Sequential(LayerNorm, Linear(hidden_dim, num_classes))`` per layer."""

    def __init__(self, seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout,
                 residual_layers: Optional[List] = None, add_input: bool = False, num_class_tokens: int = 1, num_registers: int = 0,
                 gate_type="gumbel", gate_temp: float = 1.0, gate_bias: float = 10.0, gate_threshold: float = 0.5,
                 budget_token: Union[bool, List, Literal["learnable"]] = False, num_classes: int = 10):
        super().__init__(seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout,
                         residual_layers=residual_layers, add_input=add_input, num_class_tokens=num_class_tokens,
                         num_registers=num_registers, gate_type=gate_type, gate_temp=gate_temp, gate_bias=gate_bias,
                         gate_threshold=gate_threshold, budget_token=budget_token)
        self.num_classes = num_classes
        self.early_exit_heads = nn.ModuleList([nn.Sequential(nn.LayerNorm(hidden_dim), nn.Linear(hidden_dim, num_classes))
                                               for _ in range(num_layers)])


class AViTEncoder(_EncoderBase):
    """models/adavit.py:84-219 (``layers`` is a ModuleList there)."""

    def __init__(self, seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, eps: float = 0.01,
                 gate_scale: float = 10, gate_center: float = 30):
        blocks = [AViTBlock(num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, gate_scale, gate_center)
                  for _ in range(num_layers)]
        super().__init__(seq_length, hidden_dim, dropout, blocks, module_list=True)
        self.eps, self.seq_length = eps, seq_length
        self.rho_token = self.counter_token = None
        self.halting_score_layer = []


class ViTEncoderMoE(_EncoderBase):
    """models/moevit.py:145-187."""

    def __init__(self, seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, mlp_moes=None, attn_moes=None):
        self.mlp_moes = mlp_moes or [1] * num_layers
        self.attn_moes = attn_moes or [1] * num_layers
        blocks = [ViTBlockMoE(num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, mlp_num_experts=self.mlp_moes[i],
                              attn_num_experts=self.attn_moes[i]) for i in range(num_layers)]
        mm, am = self.mlp_moes, self.attn_moes
        super().__init__(seq_length, hidden_dim, dropout, blocks)
        self.mlp_moes, self.attn_moes = mm, am


# ------------------------------------------------------------------------------ models
class _ModelBase(nn.Module):
    _family = "vit"

    def _setup(self, image_size, patch_size, hidden_dim, mlp_dim, num_heads, num_classes, dropout, attention_dropout,
               representation_size, num_class_tokens=1, num_registers=0, cls_name="class_tokens"):
        torch._assert(image_size % patch_size == 0, "Input shape indivisible by patch size!")
        self.image_size, self.patch_size = image_size, patch_size
        self.hidden_dim, self.mlp_dim = hidden_dim, mlp_dim
        self.attention_dropout, self.dropout = attention_dropout, dropout
        self.num_classes, self.representation_size, self.num_heads = num_classes, representation_size, num_heads
        self.conv_proj = nn.Conv2d(in_channels=3, out_channels=hidden_dim, kernel_size=patch_size, stride=patch_size)
        seq_length = (image_size // patch_size) ** 2
        setattr(self, cls_name, nn.Parameter(torch.zeros(1, num_class_tokens, hidden_dim)))
        seq_length += num_class_tokens
        if num_registers > 0:
            self.register_tokens = nn.Parameter(torch.zeros(1, num_registers, hidden_dim))
            seq_length += num_registers
        return seq_length

    def _finish(self, hidden_dim, num_classes):
        self.head = nn.Linear(hidden_dim, num_classes)
        nn.init.zeros_(self.head.weight)
        nn.init.zeros_(self.head.bias)
        fan_in = self.conv_proj.in_channels * self.conv_proj.kernel_size[0] * self.conv_proj.kernel_size[1]
        nn.init.trunc_normal_(self.conv_proj.weight, std=math.sqrt(1 / fan_in))
        nn.init.zeros_(self.conv_proj.bias)

    @staticmethod
    def _no_pretrained(torch_pretrained_weights, timm_pretrained_weights):
        assert not (torch_pretrained_weights and timm_pretrained_weights), \
            "You cannot load weights from both torch and timm at the same time."
        if torch_pretrained_weights is not None or timm_pretrained_weights is not None:
            raise NotImplementedError(
                "pretrained-weight download/adaptation (reference models/adapters.py) is outside the B200 hot-path scope; "
                "convert with the reference and load the resulting checkpoint with load_state_dict")

    def remove_layers(self, remove_layers: List[int]):
        """models/vit.py:301-315."""
        for i in sorted(remove_layers, reverse=True):
            del self.encoder.layers[i]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return runner.run(self, x)

    def forward_host(self, x_host: torch.Tensor, out_host: Optional[torch.Tensor] = None,
                     next_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host batch in, host logits out; H2D copies overlap compute (runner.run_host).  Besides the reference's float
        (B,3,S,S) tensors this (and forward) accepts uint8 (B,S,S,3) images as decoded: ToTensor + Normalize
        (data/imagenette.py:69-73; statistics in ``self.pk_input_norm = (mean, std)``, ImageNet by default) are then fused
        into the im2col kernel and the host->device copy is 4x smaller.  ``next_host``: the batch the NEXT call will get (what a
        prefetching loader already holds); its first chunk is copied while this batch still computes, so that call starts
        without an exposed copy.  Its leading rows must not change until then."""
        return runner.run_host(self, x_host, out_host, next_host)


class VisionTransformer(_ModelBase):
    """Drop-in for reference models/vit.py:100-315."""
    _family = "vit"

    def __init__(self, image_size: int, patch_size: int, num_layers: int, num_heads: int, hidden_dim: int, mlp_dim: int,
                 dropout: float = 0.0, attention_dropout: float = 0.0, num_classes: int = 1000,
                 representation_size: Optional[int] = None, num_registers: int = 0, num_class_tokens: int = 1,
                 torch_pretrained_weights: Optional[str] = None, timm_pretrained_weights: Optional[List] = None,
                 remove_layers: List[int] = []):
        super().__init__()
        self.num_registers, self.num_class_tokens = num_registers, num_class_tokens
        seq_length = self._setup(image_size, patch_size, hidden_dim, mlp_dim, num_heads, num_classes, dropout, attention_dropout,
                                 representation_size, num_class_tokens, num_registers)
        self.encoder = ViTEncoder(seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout)
        self.seq_length = seq_length
        self._finish(hidden_dim, num_classes)
        self._no_pretrained(torch_pretrained_weights, timm_pretrained_weights)
        if remove_layers:
            self.remove_layers(remove_layers)


class RankVisionTransformer(_ModelBase):
    """Drop-in for reference models/rankvit.py:156-339."""
    _family = "rankvit"

    def __init__(self, image_size: int, patch_size: int, num_layers: int, num_heads: int, hidden_dim: int, mlp_dim: int,
                 dropout: float = 0.0, attention_dropout: float = 0.0, num_classes: int = 1000,
                 representation_size: Optional[int] = None, num_registers: int = 0, num_class_tokens: int = 1,
                 torch_pretrained_weights: Optional[str] = None, timm_pretrained_weights: Optional[List] = None,
                 rankvit_layers: Optional[List[Union[int, float]]] = None):
        super().__init__()
        if num_registers > 0:
            raise ValueError("Registers are not supported yet for this model.")
        self.num_registers, self.num_class_tokens = num_registers, num_class_tokens
        self.rankvit_layers = rankvit_layers
        seq_length = self._setup(image_size, patch_size, hidden_dim, mlp_dim, num_heads, num_classes, dropout, attention_dropout,
                                 representation_size, num_class_tokens, 0)
        self.encoder = RankViTEncoder(seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout,
                                      rankvit_layers)
        self.seq_length = seq_length
        self._finish(hidden_dim, num_classes)
        self._no_pretrained(torch_pretrained_weights, timm_pretrained_weights)

    def set_budget(self, budget):
        """Scalar, or list indexed by layer index (models/rankvit.py:283-288)."""
        self.current_budget = budget
        for i in self.rankvit_layers:
            self.encoder.layers[i].set_budget(budget[i] if isinstance(budget, list) else budget)


class ResidualVisionTransformer(_ModelBase):
    """Drop-in for reference models/residualvit.py:352-694 (eval path)."""
    _family = "residualvit"

    def __init__(self, image_size: int, patch_size: int, num_layers: int, num_heads: int, hidden_dim: int, mlp_dim: int,
                 dropout: float = 0.0, attention_dropout: float = 0.0, num_classes: int = 1000,
                 representation_size: Optional[int] = None, num_registers: int = 0, residual_layers: Optional[List] = None,
                 add_input: bool = False, num_class_tokens: int = 1, gate_type: Literal["gumbel", "sigmoid"] = "gumbel",
                 gate_temp: float = 1.0, gate_bias: float = 10.0, gate_threshold: float = 0.5,
                 add_budget_token: Union[bool, List, Literal["learnable", "learnable_interpolate"]] = False,
                 budget_interval: Optional[List] = (0, 1), torch_pretrained_weights: Optional[str] = None,
                 timm_pretrained_weights: Optional[List] = None, remove_layers: List[int] = []):
        super().__init__()
        self.num_registers, self.num_class_tokens = num_registers, num_class_tokens
        self.add_budget_token = add_budget_token
        self.current_budget = None
        self.gate_temp, self.gate_bias, self.budget_interval = gate_temp, gate_bias, budget_interval
        self.residual_layers = residual_layers or ["attention+mlp"] * num_layers
        seq_length = self._setup(image_size, patch_size, hidden_dim, mlp_dim, num_heads, num_classes, dropout, attention_dropout,
                                 representation_size, num_class_tokens, num_registers)
        self.num_special_tokens = num_class_tokens + num_registers
        # like the reference (residualvit.py:453-468) the encoder is not told about registers
        self.encoder = ResidualViTEncoder(seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout,
                                          residual_layers=self.residual_layers, add_input=add_input, gate_type=gate_type,
                                          gate_temp=gate_temp, gate_bias=gate_bias, gate_threshold=gate_threshold,
                                          budget_token=add_budget_token)
        self.seq_length = seq_length
        if self.add_budget_token:
            self.num_budget_tokens = 1
            if self.add_budget_token == "learnable":
                self.learnable_budget_token_1 = nn.Parameter(torch.randn(1, 1, hidden_dim))
            if self.add_budget_token == "learnable_interpolate":
                self.learnable_budget_token_1 = nn.Parameter(torch.randn(1, 1, hidden_dim))
                self.learnable_budget_token_2 = nn.Parameter(torch.randn(1, 1, hidden_dim))
                self.num_budget_tokens = 2
        self._finish(hidden_dim, num_classes)
        self._no_pretrained(torch_pretrained_weights, timm_pretrained_weights)
        if remove_layers:
            self.remove_layers(remove_layers)

    def _sample_budget(self, n: int) -> torch.Tensor:
        """models/residualvit.py:541-550: the budgets of a training batch -- one draw per image from the list or from
        ``budget_interval``, or the fixed float.  Used by ``peekvit_b200.finetune.FineTuner`` (the training-mode forward)."""
        abt = self.add_budget_token
        if isinstance(abt, (list, tuple)):
            return torch.tensor([random.choice(abt) for _ in range(n)])
        if isinstance(abt, float):
            return torch.tensor(abt)
        lo, hi = self.budget_interval
        return torch.rand(n) * (hi - lo) + lo

    def set_budget(self, budget: float):
        """models/residualvit.py:619-622."""
        if self.training:
            raise ValueError("You cannot set the budget during training in this model. This model has a learnable budget so you "
                             "have to set it at the beginning of the training and then sample it during training. Use the "
                             "add_budget_token parameter to specify the budget sampling strategy.")
        self.current_budget = torch.tensor(budget, device=self.class_tokens.device)


class EEResidualVisionTransformer(_ModelBase):
    """Drop-in for reference models/eeresidualvit.py:100-363 (eval path): ResidualViT with an early-exit head after every
    layer.  ``forward`` returns a list: the L early-exit logits (each ``.squeeze()``-d like the reference, :94), then the
    final logits (:355-357)."""
    _family = "eeresidualvit"

    def __init__(self, image_size: int, patch_size: int, num_layers: int, num_heads: int, hidden_dim: int, mlp_dim: int,
                 dropout: float = 0.0, attention_dropout: float = 0.0, num_classes: int = 1000,
                 representation_size: Optional[int] = None, num_registers: int = 0, residual_layers: Optional[List] = None,
                 add_input: bool = False, num_class_tokens: int = 1, gate_type: Literal["gumbel", "sigmoid"] = "gumbel",
                 gate_temp: float = 1.0, gate_bias: float = 10.0, gate_threshold: float = 0.5,
                 add_budget_token: Union[bool, List, Literal["learnable", "learnable_interpolate"]] = False):
        super().__init__()
        self.num_registers, self.num_class_tokens = num_registers, num_class_tokens
        self.budget = add_budget_token
        self.current_budget = None
        self.gate_temp, self.gate_bias = gate_temp, gate_bias
        self.residual_layers = residual_layers or ["attention+mlp"] * num_layers
        seq_length = self._setup(image_size, patch_size, hidden_dim, mlp_dim, num_heads, num_classes, dropout, attention_dropout,
                                 representation_size, num_class_tokens, num_registers)
        self.num_special_tokens = num_class_tokens + num_registers
        # like the reference (:193-209) the encoder is told about neither class-token count nor registers
        self.encoder = EEResidualViTEncoder(seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout,
                                            residual_layers=self.residual_layers, add_input=add_input, gate_type=gate_type,
                                            gate_temp=gate_temp, gate_bias=gate_bias, gate_threshold=gate_threshold,
                                            budget_token=add_budget_token, num_classes=num_classes)
        self.seq_length = seq_length
        if self.budget:
            self.num_budget_tokens = 1
        if self.budget == "learnable" or self.budget == "learnable_interpolate":
            self.learnable_budget_token_1 = nn.Parameter(torch.randn(1, 1, hidden_dim))
            self.learnable_budget_token_2 = nn.Parameter(torch.randn(1, 1, hidden_dim))     # always both (:212-216)
        self._finish(hidden_dim, num_classes)

    def set_budget(self, budget: float):
        """models/eeresidualvit.py:361-362."""
        self.current_budget = budget

    def forward(self, x: torch.Tensor):
        out = runner.run(self, x)                     # (L + 1, B, C)
        n = out.shape[0] - 1
        return [out[i].unsqueeze(1).squeeze() for i in range(n)] + [out[n]]

    def forward_host(self, x_host, out_host=None, next_host=None):
        """Host images in, the same list as ``forward`` out (host tensors; ``out_host``: optional pinned (L + 1, B, C) buffer)."""
        out = runner.run_host(self, x_host, out_host, next_host)   # (L + 1, B, C)
        n = out.shape[0] - 1
        return [out[i].unsqueeze(1).squeeze() for i in range(n)] + [out[n]]


class AdaptiveVisionTransformer(_ModelBase):
    """Drop-in for reference models/adavit.py:225-433."""
    _family = "adavit"

    def __init__(self, image_size: int, patch_size: int, num_layers: int, num_heads: int, hidden_dim: int, mlp_dim: int,
                 dropout: float = 0.0, attention_dropout: float = 0.0, num_classes: int = 1000,
                 representation_size: Optional[int] = None, num_registers: int = 0, num_class_tokens: int = 1,
                 eps: float = 0.01, gate_scale: float = 10, gate_center: float = 30,
                 torch_pretrained_weights: Optional[str] = None, timm_pretrained_weights: Optional[List] = None):
        super().__init__()
        self.num_registers, self.num_class_tokens = num_registers, num_class_tokens
        self.num_layers, self.eps, self.gate_scale, self.gate_center = num_layers, eps, gate_scale, gate_center
        seq_length = self._setup(image_size, patch_size, hidden_dim, mlp_dim, num_heads, num_classes, dropout, attention_dropout,
                                 representation_size, num_class_tokens, num_registers)
        self.encoder = AViTEncoder(seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, eps,
                                   gate_scale, gate_center)
        self.seq_length = seq_length
        self._finish(hidden_dim, num_classes)
        self._no_pretrained(torch_pretrained_weights, timm_pretrained_weights)


class VisionTransformerMoE(_ModelBase):
    """Drop-in for reference models/moevit.py:191-315 (parameter ``class_token``, singular)."""
    _family = "moevit"

    def __init__(self, image_size: int, patch_size: int, num_layers: int, num_heads: int, hidden_dim: int, mlp_dim: int,
                 dropout: float = 0.0, attention_dropout: float = 0.0, num_classes: int = 1000,
                 representation_size: Optional[int] = None, mlp_moes: List = None, attn_moes: List = None):
        super().__init__()
        self.mlp_moes = mlp_moes or [1] * num_layers
        self.attn_moes = attn_moes or [1] * num_layers
        seq_length = self._setup(image_size, patch_size, hidden_dim, mlp_dim, num_heads, num_classes, dropout, attention_dropout,
                                 representation_size, 1, 0, cls_name="class_token")
        self.encoder = ViTEncoderMoE(seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, mlp_moes,
                                     attn_moes)
        self.seq_length = seq_length
        self._finish(hidden_dim, num_classes)
