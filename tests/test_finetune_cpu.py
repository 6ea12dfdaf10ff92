"""Host logic of the fine-tuning path without a GPU: the freezing rule of the reference
(models/topology.py:128-158), which regimes the FineTuner accepts, and the gradient all-reduce on two gloo ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from peekvit_b200 import finetune
from peekvit_b200.models import build_model

_BASE = dict(image_size=32, patch_size=8, num_layers=2, num_heads=2, hidden_dim=64, mlp_dim=128, num_classes=10)


def test_train_only_these_params_matches_the_reference_rule():
    m = build_model("vit", dict(_BASE, num_class_tokens=2))
    names = finetune.train_only_these_params(m)                       # ['gate', 'class', 'head', 'threshold', 'budget']
    assert sorted(names) == ["class_tokens", "head.bias", "head.weight"]
    assert all(p.requires_grad == (n in names) for n, p in m.named_parameters())
    r = build_model("residualvit", dict(_BASE, gate_type="sigmoid", gate_bias=0.0, add_budget_token="learnable",
                                        residual_layers=["attention+mlp"] * 2))
    names = finetune.train_only_these_params(r)
    assert "learnable_budget_token_1" in names and "encoder.layers.0.residual_gate.projection.weight" in names
    assert "encoder.layers.1.budget_token_gate.bias" in names and "encoder.layers.0.mlp.fc1.weight" not in names


def test_finetuner_accepts_the_dense_regime_only():
    ft = finetune.FineTuner(build_model("vit", _BASE))
    assert sorted(ft.params) == ["class_tokens", "head.bias", "head.weight"]
    assert sorted(finetune.FineTuner(build_model("RankVisionTransformer", dict(_BASE, rankvit_layers=[1]))).params) == sorted(ft.params)
    with pytest.raises(NotImplementedError):                          # gates / budget tokens: training-mode forward not built
        finetune.FineTuner(build_model("residualvit", dict(_BASE, gate_type="sigmoid", gate_bias=0.0, add_budget_token="learnable",
                                                           residual_layers=["attention+mlp"] * 2)))
    with pytest.raises(NotImplementedError):                          # a trainable backbone parameter has no weight-gradient kernel
        finetune.FineTuner(build_model("vit", _BASE), train_words=("head", "fc1"))
    with pytest.raises(NotImplementedError):
        finetune.FineTuner(build_model("vit", dict(_BASE, dropout=0.1)))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ft.forward_backward(torch.randn(2, 3, 32, 32), torch.zeros(2, dtype=torch.long))


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(1, 2, 4))]
        per_rank = [[torch.randn_like(p) for p in params] for _ in range(world)]      # every rank draws all ranks' gradients
        for p, g in zip(params, per_rank[rank]):
            p.grad = g.clone()
        assert finetune.all_reduce_mean_(params) == world
        for i, p in enumerate(params):
            mean = sum(per_rank[r][i] for r in range(world)) / world
            assert torch.allclose(p.grad, mean, atol=1e-7)
        bucket = finetune.flatten_grads(params)
        assert bucket.numel() == 15 + 7 + 8
    finally:
        dist.destroy_process_group()


def test_gradient_all_reduce_two_ranks():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)


def test_all_reduce_is_a_no_op_without_a_process_group():
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.arange(4.0)
    assert finetune.all_reduce_mean_([p]) == 1 and torch.equal(p.grad, torch.arange(4.0))
