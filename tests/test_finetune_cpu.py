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
    # the gate regime of ResidualViT as shipped (sigmoid gates, learnable budget token): gates, budget-token gates, the
    # learnable budget token, class token and head train; other configurations have no backward
    res = finetune.FineTuner(build_model("residualvit", dict(_BASE, gate_type="sigmoid", gate_bias=0.0, add_budget_token="learnable",
                                                            residual_layers=["attention+mlp", "none"])))
    assert sorted(res.params) == sorted(
        ["class_tokens", "head.bias", "head.weight", "learnable_budget_token_1"]
        + [f"encoder.layers.0.residual_gate.projection.{k}" for k in ("weight", "bias")]
        + [f"encoder.layers.{i}.budget_token_gate.{k}" for i in (0, 1) for k in ("weight", "bias")])
    with pytest.raises(NotImplementedError):
        finetune.FineTuner(build_model("residualvit", dict(_BASE, gate_type="gumbel", gate_bias=0.0, add_budget_token="learnable",
                                                           residual_layers=["attention+mlp"] * 2)))
    with pytest.raises(NotImplementedError):
        finetune.FineTuner(build_model("residualvit", dict(_BASE, gate_type="sigmoid", gate_bias=0.0, add_budget_token=[0.3, 0.6],
                                                           residual_layers=["attention+mlp"] * 2)))
    with pytest.raises(NotImplementedError):
        finetune.FineTuner(build_model("adavit", _BASE))
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


def test_oracle_reproduces_the_reference_gate_regime_gradients():
    """tests/golden/finetune_residual_learnable.npz holds logits, loss, masks and the 20 parameter gradients of the REFERENCE
    ResidualViT in train() mode (make_finetune_residual.py).  The oracle restatement, differentiated by torch autograd with the
    same per-image budgets and the same regulariser, must reproduce them: this pins the checker of the GPU gradient tests."""
    import os
    import numpy as np
    from golden_cases import CASES, build_case
    from oracle import peekvit_oracle as po
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "finetune_residual_learnable.npz"))
    case = CASES["residual_learnable_cal04"]
    sd, images = build_case(case)
    B = images.shape[0]
    labels, budgets = torch.from_numpy(fx["labels"]), torch.from_numpy(fx["budgets"])
    names = [k[5:] for k in fx.files if k.startswith("grad.")]
    assert len(names) == 20 and "learnable_budget_token_1" in names and "encoder.layers.3.budget_token_gate.bias" in names
    sdg = {k: v.clone() for k, v in sd.items()}
    for n in names:
        sdg[n].requires_grad_(True)
    logits, aux = po.residualvit_forward(sdg, case["cfg"], images, budgets.view(B, 1, 1))
    masks = [aux["masks"][i] for i in sorted(aux["masks"])]
    sp = torch.stack([m.mean(dim=(1, 2)) for m in masks]).mean()
    reg = ((sp - budgets) ** 2).sum().mul(2 - budgets).mean()          # utils/losses.py:111-142, per_layer=False, strict=True
    loss = torch.nn.functional.cross_entropy(logits, labels) + 0.5 * reg
    loss.backward()
    assert abs(loss.item() - float(fx["loss"])) < 1e-5
    assert (logits.detach() - torch.from_numpy(fx["logits"])).abs().max().item() < 1e-5
    for i, m in enumerate(masks):
        assert (m.detach() - torch.from_numpy(fx[f"mask.{i}"])).abs().max().item() < 1e-6
    for n in names:
        want = torch.from_numpy(fx["grad." + n])
        assert ((sdg[n].grad - want).abs().max() / want.abs().max()).item() < 1e-4, n


@pytest.mark.parametrize("name", ["vit_d128_regs", "rankvit_b05"])
def test_oracle_reproduces_the_reference_class_token_and_head_gradients(name):
    """tests/golden/finetune_<case>.npz: logits, loss and the class_tokens / head gradients of the REFERENCE model in train()
    mode under train_only_these_params (make_finetune_vit.py).  The oracle under torch autograd must reproduce them."""
    import os
    import numpy as np
    from golden_cases import CASES, build_case
    from oracle import peekvit_oracle as po
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", f"finetune_{name}.npz"))
    case = CASES[name]
    sd, images = build_case(case)
    labels = torch.from_numpy(fx["labels"])
    names = ["class_tokens", "head.weight", "head.bias"]
    sdg = {k: v.clone() for k, v in sd.items()}
    for n in names:
        sdg[n].requires_grad_(True)
    if case["family"] == "vit":
        logits, _ = po.vit_forward(sdg, case["cfg"], images)
    else:
        logits, _ = po.rankvit_forward(sdg, case["cfg"], images, case["budget"])
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    assert abs(loss.item() - float(fx["loss"])) < 1e-5
    assert (logits.detach() - torch.from_numpy(fx["logits"])).abs().max().item() < 1e-5
    for n in names:
        want = torch.from_numpy(fx["grad." + n])
        assert ((sdg[n].grad.view_as(want) - want).abs().max() / want.abs().max()).item() < 1e-4, n
