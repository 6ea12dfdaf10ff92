"""Fine-tuning path on the GPU: every backward kernel against torch autograd on the same values, then the
class-token / head gradients of whole models against autograd through the oracle's fp32 forward (reference semantics:
train/train.py:105-113 with models/topology.py:128-158), and a few optimiser steps."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("rows,D", [(300, 384), (77, 768), (5, 64), (1000, 192)])
def test_layernorm_bwd_against_autograd(rows, D):
    from peekvit_b200 import ops
    x = torch.randn(rows, D, device=DEV) * 2 + 0.3
    dy = torch.randn(rows, D, device=DEV)
    gamma = 1 + 0.1 * torch.randn(D, device=DEV)
    xr = x.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (D,), gamma, torch.zeros(D, device=DEV), 1e-6).backward(dy)
    base = torch.randn(rows, D, device=DEV)
    dx = base.clone()
    ops.layernorm_bwd(x, dy, gamma, 1e-6, dx, rows)                        # accumulates
    assert _rel(dx - base, xr.grad) < 1e-5
    # class-row form: rows picked through an index, several rows sharing one dy row, overwrite
    idx = torch.tensor([3, 1, rows - 1, 0], device=DEV, dtype=torch.int32)
    dy2 = torch.randn(2, D, device=DEV)
    out = torch.full((rows, D), 7.0, device=DEV)
    ops.layernorm_bwd(x, dy2, gamma, 1e-6, out, 4, row_index=idx, dy_div=2, accumulate=False)
    xr = x.clone().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xr, (D,), gamma, torch.zeros(D, device=DEV), 1e-6)
    (y[idx.long()] * dy2.repeat_interleave(2, 0)).sum().backward()
    assert _rel(out[idx.long()], xr.grad[idx.long()]) < 1e-5
    keep = torch.ones(rows, dtype=torch.bool, device=DEV)
    keep[idx.long()] = False
    assert bool((out[keep] == 7.0).all())
    assert ops.device_flag() == 0


def test_gelu_forward_backward_against_autograd():
    from peekvit_b200 import ops
    h = (torch.randn(64, 1536, device=DEV) * 2).to(torch.bfloat16)
    dy = torch.randn(64, 1536, device=DEV).to(torch.bfloat16)
    y = ops.gelu_bf16(h, torch.empty_like(h))
    assert _rel(y, torch.nn.functional.gelu(h.float())) < 4e-3                     # bf16 rounding of the result
    hr = h.float().requires_grad_(True)
    torch.nn.functional.gelu(hr).backward(dy.float())
    dx = ops.gelu_bwd_bf16(h, dy, torch.empty_like(h))
    assert _rel(dx, hr.grad) < 4e-3
    out = ops.cast_bf16(hr.detach(), torch.empty_like(h))
    assert torch.equal(out, h)


@pytest.mark.parametrize("B,H,dh,n", [(3, 6, 64, 197), (2, 2, 64, 65), (4, 3, 32, 50), (1, 1, 64, 1), (2, 12, 64, 256), (2, 8, 32, 17)])
def test_attention_bwd_against_autograd(B, H, dh, n):
    from peekvit_b200 import ops
    D = H * dh
    qkv = (torch.randn(B * n, 3 * D, device=DEV) * 0.8).to(torch.bfloat16)
    dout = torch.randn(B * n, D, device=DEV).to(torch.bfloat16)
    x = qkv.float().view(B, n, 3, H, dh).requires_grad_(True)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    o = (torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), -1) @ v).transpose(1, 2).reshape(B * n, D)
    o.backward(dout.float())
    ref = x.grad.reshape(B * n, 3 * D)
    dqkv = torch.full((B * n, 3 * D), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attention_bwd(qkv, o.detach().to(torch.bfloat16), dout, dqkv, B, H, dh, n)
    assert ops.device_flag() == 0
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        if n == 1 and name != "dv":
            # a single key: the softmax is the constant 1 and dq = dk = 0 exactly; here dS is the bf16 rounding residue of O
            assert dqkv[:, sl].float().abs().max().item() < 1e-2 * ref.abs().max().item(), name
            continue
        assert _rel(dqkv[:, sl], ref[:, sl]) < 1e-2, name                               # bf16 rounding of the stored O and of the result


def test_softmax_xent_and_head_bwd_against_autograd():
    from peekvit_b200 import ops
    B, C, D = 37, 1000, 384
    feat = torch.randn(B, D, device=DEV)
    W = (torch.randn(C, D, device=DEV) / math.sqrt(D)).requires_grad_(True)
    bias = (torch.randn(C, device=DEV) * 0.1).requires_grad_(True)
    fr = feat.clone().requires_grad_(True)
    labels = torch.randint(0, C, (B,), device=DEV)
    logits = fr @ W.t() + bias
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    loss_sum = torch.zeros(1, device=DEV)
    correct = torch.zeros(1, device=DEV, dtype=torch.int32)
    dlogits = torch.empty(B, C, device=DEV)
    ops.softmax_xent(logits.detach(), labels, 1.0 / B, loss_sum, dlogits, correct)
    assert abs(loss_sum.item() - loss.item()) < 1e-5 * abs(loss.item())
    assert correct.item() == int((logits.argmax(1) == labels).sum())
    dW, db, dfeat = torch.zeros(C, D, device=DEV), torch.zeros(C, device=DEV), torch.empty(B, D, device=DEV)
    ops.head_bwd(dlogits, feat, W.detach(), dW, db, dfeat)
    assert _rel(dW, W.grad) < 1e-5 and _rel(db, bias.grad) < 1e-5 and _rel(dfeat, fr.grad) < 1e-5
    ops.head_bwd(dlogits, feat, W.detach(), dW, db, dfeat)                         # weight / bias gradients accumulate
    assert _rel(dW, 2 * W.grad) < 1e-5
    x = torch.randn(5 * 9, 64, device=DEV)
    out = torch.ones(2, 64, device=DEV)
    ops.sum_token_rows(x, 5, 9, 3, 2, out)
    assert torch.allclose(out, 1 + x.view(5, 9, 64)[:, 3:5].sum(0), atol=1e-5)
    assert ops.device_flag() == 0


def _autograd_reference(sd, cfg, images, labels):
    """Gradients of the fine-tuning regime through the oracle's fp32 forward with torch autograd (on the GPU: test
    infrastructure).  Returns (loss, {name: grad})."""
    from oracle import peekvit_oracle as po
    names = ("class_tokens", "head.weight", "head.bias")
    sdg = {k: v.to(DEV) for k, v in sd.items()}
    for n in names:
        sdg[n] = sdg[n].clone().requires_grad_(True)
    logits, _ = po.vit_forward(sdg, cfg, images.to(DEV))
    loss = torch.nn.functional.cross_entropy(logits, labels.to(DEV))
    loss.backward()
    return loss.detach(), logits.detach(), {n: sdg[n].grad for n in names}


@pytest.mark.parametrize("cfg,batch,mb", [
    (dict(image_size=64, patch_size=8, num_layers=4, num_heads=2, hidden_dim=128, mlp_dim=256, num_classes=10), 6, 4),
    (dict(image_size=48, patch_size=16, num_layers=2, num_heads=2, hidden_dim=128, mlp_dim=256, num_classes=7, num_registers=2,
          num_class_tokens=2), 5, 128),
    (dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000), 6, 4),
])
def test_class_token_and_head_gradients_match_autograd(cfg, batch, mb):
    """loss.backward() of the reference regime (train/train.py:105-113): gradients of class_tokens / head.weight / head.bias
    from the CUDA path (bf16 operands, bf16 activation gradients between the dX GEMMs) against fp32 autograd through the
    oracle; the third case is the ViT-B/16 shape (tcgen05 attention + CTA-pair GEMMs in the forward and in the dX GEMMs),
    split into micro-batches whose gradients accumulate."""
    from oracle import weights as ow
    from peekvit_b200 import ops
    from peekvit_b200.finetune import FineTuner
    from peekvit_b200.models import VisionTransformer
    sd = ow.make_state_dict("vit", cfg, seed=77)
    images = ow.synthetic_images(batch, cfg["image_size"], seed=5)
    labels = torch.randint(0, cfg["num_classes"], (batch,), generator=torch.Generator().manual_seed(1))
    loss_ref, logits_ref, grads = _autograd_reference(sd, cfg, images, labels)
    model = VisionTransformer(**cfg)
    model.load_state_dict(sd)
    model = model.to(DEV).train()
    ft = FineTuner(model, micro_batch=mb)
    loss, logits = ft.forward_backward(images.to(DEV), labels.to(DEV))
    assert ops.device_flag() == 0
    assert _rel(logits, logits_ref) < 1e-2
    assert abs(loss.item() - loss_ref.item()) < 1e-2 * abs(loss_ref.item())
    for n, p in ft.params.items():
        err = _rel(p.grad, grads[n].view_as(p.grad))
        print(f"{n}: rel err {err:.2e} (max |grad| {grads[n].abs().max().item():.2e})")
        assert err < 3e-2, n
    frozen = [n for n, p in model.named_parameters() if not p.requires_grad]
    assert "encoder.layers.0.mlp.fc1.weight" in frozen and all(model.get_parameter(n).grad is None for n in frozen)
    # a second call accumulates like autograd
    ft.forward_backward(images.to(DEV), labels.to(DEV))
    assert _rel(model.head.bias.grad, 2 * grads["head.bias"]) < 3e-2


def test_a_few_sgd_steps_reduce_the_loss_and_eval_follows_the_new_weights():
    from oracle import weights as ow
    from peekvit_b200.finetune import FineTuner
    from peekvit_b200.models import VisionTransformer
    cfg = dict(image_size=64, patch_size=8, num_layers=3, num_heads=2, hidden_dim=128, mlp_dim=256, num_classes=10)
    model = VisionTransformer(**cfg)
    model.load_state_dict(ow.make_state_dict("vit", cfg, seed=3))
    model = model.to(DEV)
    images = ow.synthetic_images(16, 64, seed=9).to(DEV)
    labels = torch.arange(16, device=DEV) % 10
    model.eval()
    before = model(images)
    model.train()
    ft = FineTuner(model)
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=0.05)
    losses = []
    for _ in range(12):
        opt.zero_grad()
        loss, _ = ft.forward_backward(images, labels)
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < 0.7 * losses[0], losses
    model.eval()
    after = model(images)                                                   # inference path sees the updated class tokens / head
    assert (after - before).abs().max().item() > 1e-2
    assert torch.nn.functional.cross_entropy(after, labels).item() < torch.nn.functional.cross_entropy(before, labels).item()


@pytest.mark.parametrize("budget", [0.5, [1, 0.5, 1, 0.25]])
def test_rankvit_gradients_match_autograd_given_identical_selections(budget):
    """RankVisionTransformer in the same regime: the blocks rank and drop tokens in training as in eval (rankvit.py:85-88),
    the backward scatters the gradient rows back through the gather.  top-k is discontinuous in the (bf16-level) scores, so
    the autograd reference runs the oracle with the selections the CUDA path made (forced_kept), like the inference tests."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200 import ops
    from peekvit_b200.finetune import FineTuner
    from peekvit_b200.models import RankVisionTransformer
    cfg = dict(image_size=64, patch_size=8, num_layers=4, num_heads=2, hidden_dim=128, mlp_dim=256, num_classes=10, rankvit_layers=[1, 3])
    sd = ow.make_state_dict("rankvit", cfg, seed=13)
    images = ow.synthetic_images(5, 64, seed=23)
    labels = torch.tensor([1, 7, 3, 3, 9])
    model = RankVisionTransformer(**cfg)
    model.load_state_dict(sd)
    model = model.to(DEV).train()
    model.set_budget(budget)
    ft = FineTuner(model)
    loss, logits = ft.forward_backward(images.to(DEV), labels.to(DEV))
    assert ops.device_flag() == 0 and sorted(ft.last_kept) == [1, 3]
    assert ft.last_kept[1].shape == (5, 32) and ft.last_kept[3].shape == (5, 16 if budget == 0.5 else 8)
    names = ("class_tokens", "head.weight", "head.bias")
    sdg = {k: v.to(DEV) for k, v in sd.items()}
    for n in names:
        sdg[n] = sdg[n].clone().requires_grad_(True)
    ref_logits, _ = po.rankvit_forward(sdg, cfg, images.to(DEV), budget, forced_kept={l: k.long() for l, k in ft.last_kept.items()})
    ref_loss = torch.nn.functional.cross_entropy(ref_logits, labels.to(DEV))
    ref_loss.backward()
    assert _rel(logits, ref_logits.detach()) < 1e-2 and abs(loss.item() - ref_loss.item()) < 1e-2 * ref_loss.item()
    for n, p in ft.params.items():
        err = _rel(p.grad, sdg[n].grad.view_as(p.grad))
        print(f"rankvit {n}: rel err {err:.2e}")
        assert err < 3e-2, n


def test_scatter_rows_is_the_adjoint_of_gather_rows():
    from peekvit_b200 import ops
    B, seq, k, D = 3, 9, 4, 64
    x = torch.randn(B * seq, D, device=DEV)
    kept = torch.stack([torch.randperm(seq - 1, device=DEV)[:k] for _ in range(B)]).to(torch.int32)
    y = ops.gather_rows(x, kept, B, seq)
    g = torch.randn_like(y)
    back = ops.scatter_rows(g, torch.zeros_like(x), kept, B, seq)
    assert abs(float((y * g).sum()) - float((x * back).sum())) < 1e-3            # <gather(x), g> == <x, scatter(g)>
    assert int((back.abs().sum(1) > 0).sum()) == B * (k + 1)
