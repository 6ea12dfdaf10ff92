"""Host-side policy of the runner without a GPU: micro-batch sizes per family and path, the budget forms the compacted
ResidualViT forward accepts, the chunk schedule of the host-resident path."""
import pytest
import torch

from peekvit_b200 import engine, runner
from peekvit_b200.models import build_model

_BASE = dict(image_size=32, patch_size=8, num_layers=2, num_heads=2, hidden_dim=64, mlp_dim=128, num_classes=10)


def test_micro_batch_defaults_per_family_and_path():
    vit = build_model("vit", _BASE)
    res = build_model("residualvit", dict(_BASE, gate_type="sigmoid", gate_bias=0.0, add_budget_token="learnable",
                                          residual_layers=["attention+mlp"] * 2))
    avit = build_model("adavit", _BASE)
    moe = build_model("vitmoe", dict(_BASE, mlp_moes=[2, 2]))
    rank = build_model("RankVisionTransformer", dict(_BASE, rankvit_layers=[1]))
    assert runner._micro_batch(vit, 4096) == runner.DEFAULT_MICRO_BATCH == runner._micro_batch(rank, 4096)
    for m in (res, avit, moe):          # compacted / routed rows: many small launches per layer -> larger micro-batches on the device
        assert runner._micro_batch(m, 4096) == runner.SPARSE_MICRO_BATCH
        assert runner._micro_batch(m, 4096, host=True) == runner.DEFAULT_MICRO_BATCH        # H2D chunks overlap compute
    res.pk_micro_batch = 96
    assert runner._micro_batch(res, 4096) == 96 == runner._micro_batch(res, 4096, host=True)
    avit.pk_precision = "fp32"          # split operands are 6x wider
    assert runner._micro_batch(avit, 4096) == runner.EXACT_MICRO_BATCH
    # a fixed-float budget token thresholds on the mean over the WHOLE batch (residualvit.py:208): never split
    fixed = build_model("residualvit", dict(_BASE, gate_type="sigmoid", gate_bias=0.0, add_budget_token=0.5,
                                            residual_layers=["attention+mlp"] * 2))
    assert runner._micro_batch(fixed, 5000) == 5000


def test_light_parameters_of_the_fine_tuning_regimes():
    assert all(engine.is_light_param(n) for n in (
        "class_tokens", "class_token", "head.weight", "head.bias", "learnable_budget_token_1", "learnable_budget_token_2",
        "encoder.layers.3.residual_gate.projection.weight", "encoder.layers.11.budget_token_gate.bias"))
    assert not any(engine.is_light_param(n) for n in (
        "encoder.layers.0.mlp.fc1.weight", "conv_proj.weight", "encoder.pos_embedding", "encoder.ln.weight",
        "encoder.layers.0.self_attention.self_attention.in_proj_weight", "encoder.layers.0.residual_gate.threshold"))


def test_per_image_budgets_are_rejected_by_the_inference_path():
    """A training step leaves one sampled budget per image in ``current_budget`` (residualvit.py:565-566); the compacted
    forward takes one budget per batch, so evaluation has to call set_budget first -- checked before any device work."""
    res = build_model("residualvit", dict(_BASE, gate_type="sigmoid", gate_bias=0.0, add_budget_token="learnable",
                                          residual_layers=["attention+mlp"] * 2)).eval()
    b = res._sample_budget(7)
    assert b.shape == (7,) and bool(((b >= 0) & (b <= 1)).all())
    lst = build_model("residualvit", dict(_BASE, gate_type="sigmoid", gate_bias=0.0, add_budget_token=[0.25, 0.75],
                                          residual_layers=["attention+mlp"] * 2))
    assert set(lst._sample_budget(50).tolist()) <= {0.25, 0.75}
    assert runner.HOST_FIRST_SPLIT == (0.25, 0.75) and abs(sum(runner.HOST_FIRST_SPLIT) - 1.0) < 1e-9
