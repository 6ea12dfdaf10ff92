"""Generate tests/golden/adapter_keys.json from the imported reference (authoring container only): the key renames the
reference's adapters (models/adapters.py:75-166) apply to the parameter names of a torchvision ViT and of a timm ViT."""
import json, os, sys, tempfile
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
tmp = tempfile.mkdtemp(prefix="peekvit_ref_")
os.symlink("/root/reference", os.path.join(tmp, "peekvit"))
sys.path.insert(0, tmp)
from peekvit.models.adapters import adapt_timm_state_dict, adapt_torch_state_dict  # noqa: E402


def names_torchvision(layers=12):
    ks = ["class_token", "conv_proj.weight", "conv_proj.bias", "encoder.pos_embedding"]
    for i in range(layers):
        p = f"encoder.layers.encoder_layer_{i}."
        ks += [p + s for s in ("ln_1.weight", "ln_1.bias", "self_attention.in_proj_weight", "self_attention.in_proj_bias",
                               "self_attention.out_proj.weight", "self_attention.out_proj.bias", "ln_2.weight", "ln_2.bias",
                               "mlp.0.weight", "mlp.0.bias", "mlp.3.weight", "mlp.3.bias",
                               "mlp.linear_1.weight", "mlp.linear_1.bias", "mlp.linear_2.weight", "mlp.linear_2.bias")]
    return ks + ["encoder.ln.weight", "encoder.ln.bias", "heads.head.weight", "heads.head.bias"]


def names_timm(layers=12):
    ks = ["cls_token", "pos_embed", "patch_embed.proj.weight", "patch_embed.proj.bias"]
    for i in range(layers):
        p = f"blocks.{i}."
        ks += [p + s for s in ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
                               "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias")]
    return ks + ["norm.weight", "norm.bias", "head.weight", "head.bias"]


def run(fn, names, head):
    res = {}
    for k in names:                      # one key at a time (old- and new-style torchvision MLP names map to the same target)
        t = torch.zeros(1000, 8) if k.endswith("head.weight") else torch.zeros(1)
        sd = {head + ".weight": torch.zeros(1000, 8), head + ".bias": torch.zeros(1000)}
        sd[k] = t
        out = fn(sd, 1000)
        res[k] = next(nk for nk, v in out.items() if v is t)
    return res


json.dump({"torchvision": run(adapt_torch_state_dict, names_torchvision(), "heads.head"),
           "timm": run(adapt_timm_state_dict, names_timm(), "head")},
          open(os.path.join(HERE, "adapter_keys.json"), "w"), indent=0, sort_keys=True)
print("ok")
