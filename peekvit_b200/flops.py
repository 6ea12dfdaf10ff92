"""Analytic FLOP accounting from token counts: replaces the reference's second, hooked forward
(``validate/test.py:136-156`` -> ``utils/flops_count.py:27-180``) by the closed form

    FLOPs = 2*P*Kp*D + sum_l [ 2*n_l*D*3D + 4*n_l^2*D + 2*n_l*D^2 + 4*n_l*D*F ] + 2*D*C        (2 FLOP per MAC)

where ``n_l`` is the number of tokens that enter layer ``l`` (kept tokens for the budgeted models; fractional averages over
a batch are fine).  The kept-token counts come out of the forward itself (``aux['seq_lens']`` for RankViT, the per-layer row
counts for ResidualViT / A-ViT), so no extra pass over the model is needed.
"""
from __future__ import annotations

from typing import Optional, Sequence


def gflops_per_image(image_size: int, patch_size: int, hidden_dim: int, mlp_dim: int, num_layers: int, num_classes: int,
                     tokens_per_layer: Optional[Sequence[float]] = None, extra_tokens: int = 1) -> float:
    """GFLOP of one forward.  ``tokens_per_layer[l]`` = tokens entering layer l (default: all patches + ``extra_tokens``)."""
    P = (image_size // patch_size) ** 2
    D, F = hidden_dim, mlp_dim
    if tokens_per_layer is None:
        tokens_per_layer = [P + extra_tokens] * num_layers
    if len(tokens_per_layer) != num_layers:
        raise ValueError(f"need {num_layers} token counts, got {len(tokens_per_layer)}")
    total = 2.0 * P * 3 * patch_size ** 2 * D + 2.0 * D * num_classes
    for n in tokens_per_layer:
        total += 2.0 * n * D * 3 * D + 4.0 * n * n * D + 2.0 * n * D * D + 4.0 * n * D * F
    return total / 1e9


def model_gflops_per_image(model, tokens_per_layer: Optional[Sequence[float]] = None) -> float:
    """Same, reading the shape from a (reference or drop-in) model instance."""
    blk = next(b for b in model.encoder.layers if hasattr(b, "mlp"))       # a NoiseBlock may sit at index 0
    mlp = blk.mlp.experts[0] if hasattr(blk.mlp, "experts") else blk.mlp
    extra = int(getattr(model, "num_class_tokens", 1) or 1) + int(getattr(model, "num_registers", 0) or 0)
    if getattr(model, "add_budget_token", False):
        extra += 1
    return gflops_per_image(model.image_size, model.patch_size, model.hidden_dim, mlp.fc1.out_features, len(model.encoder.layers),
                            model.num_classes, tokens_per_layer, extra_tokens=extra)
