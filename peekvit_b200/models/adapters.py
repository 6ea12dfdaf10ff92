"""Checkpoint adapters (drop-ins for reference models/adapters.py): rename torchvision / timm ViT state-dict keys to the
peekvit checkpoint contract and warm-start the residual models from a plain ViT checkpoint.  Pure host-side
key maps — tensors are passed through untouched, the CUDA path prepacks them on first use (runner.packed)."""
from __future__ import annotations

import re
from typing import Dict, Optional

import torch

from .core import EEResidualVisionTransformer, ResidualVisionTransformer

# (pattern, replacement) applied in order to every torchvision key (reference adapters.py:75-116)
_TORCHVISION_RULES = (
    (r"mlp\.0\b", "mlp.fc1"), (r"mlp\.3\b", "mlp.fc2"), (r"heads\.head", "head"),
    (r"mlp\.linear_1", "mlp.fc1"), (r"mlp\.linear_2", "mlp.fc2"),
)
# timm names (reference adapters.py:119-166)
_TIMM_RULES = (
    ("norm1", "ln_1"), ("norm2", "ln_2"),
    ("attn.qkv.bias", "self_attention.self_attention.in_proj_bias"),
    ("attn.qkv.weight", "self_attention.self_attention.in_proj_weight"),
    ("attn.proj.bias", "self_attention.self_attention.out_proj.bias"),
    ("attn.proj.weight", "self_attention.self_attention.out_proj.weight"),
    ("patch_embed.proj.bias", "conv_proj.bias"), ("patch_embed.proj.weight", "conv_proj.weight"),
    ("cls_token", "class_tokens"), ("pos_embed", "encoder.pos_embedding"),
    ("norm.weight", "encoder.ln.weight"), ("norm.bias", "encoder.ln.bias"),
)


def _fit_head(sd: Dict[str, torch.Tensor], num_classes: int) -> Dict[str, torch.Tensor]:
    """A head for a different label set starts from zeros, as in the reference (adapters.py:107-114)."""
    rows, cols = sd["head.weight"].shape
    if rows != num_classes:
        print("Loading weights for a different number of classes. Replacing head with random weights. You should fine-tune the model.")
        sd["head.weight"] = torch.zeros((num_classes, cols))
        sd["head.bias"] = torch.zeros(num_classes)
    return sd


def torchvision_key(name: str) -> str:
    for pat, rep in _TORCHVISION_RULES:
        name = re.sub(pat, rep, name)
    if name.count("self_attention") == 1:
        name = name.replace("self_attention", "self_attention.self_attention")
    if name == "class_token":
        return "class_tokens"
    return re.sub(r"encoder_layer_(\d)", r"\1", name)


def timm_key(name: str) -> str:
    for old, new in _TIMM_RULES:
        name = name.replace(old, new)
    return re.sub(r"blocks.(\d+)", r"encoder.layers.\1", name)


def adapt_torch_state_dict(torch_state_dict, num_classes: int):
    """torchvision ``vit_*`` state dict -> peekvit VisionTransformer keys (reference adapters.py:75-116)."""
    return _fit_head({torchvision_key(k): v for k, v in torch_state_dict.items()}, num_classes)


def adapt_timm_state_dict(timm_state_dict, num_classes: int):
    """timm ``vit_*`` / ``deit_*`` state dict -> peekvit VisionTransformer keys (reference adapters.py:119-166)."""
    return _fit_head({timm_key(k): v for k, v in timm_state_dict.items()}, num_classes)


@torch.no_grad()
def from_vit_to_residual_vit(vit_checkpoint, model_args: Optional[dict] = None):
    """ViT checkpoint file -> ResidualVisionTransformer with the shared weights copied and the gates freshly initialised
    (reference adapters.py:8-38)."""
    state = torch.load(vit_checkpoint)
    print("Loading weights from class: ", state["model_class"])
    model_args = model_args if model_args is not None else state["model_args"]
    model = ResidualVisionTransformer(**model_args)
    res = model.load_state_dict(state["state_dict"], strict=False)
    print("Some parameters are not present in the checkpoint and will be randomly initialized: ", res[0])
    return model


@torch.no_grad()
def from_vit_to_eeresidual_vit(vit_checkpoint, residual_vit_args: Optional[dict] = None):
    """ViT checkpoint file -> (EEResidualVisionTransformer, merged model args) (reference adapters.py:42-72)."""
    state = torch.load(vit_checkpoint)
    print("Loading weights from class: ", state["model_class"])
    model_args = state["model_args"]
    model = EEResidualVisionTransformer(**model_args, **residual_vit_args)
    res = model.load_state_dict(state["state_dict"], strict=False)
    print("Some parameters are not present in the checkpoint and will be randomly initialized: ", res[0])
    model_args.update(residual_vit_args)
    return model, model_args
