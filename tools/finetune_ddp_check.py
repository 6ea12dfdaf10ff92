"""Gradient all-reduce of the fine-tuning path over NCCL: torchrun --nproc-per-node N tools/finetune_ddp_check.py [residualvit]
(default: dense ViT-S, class token + head; ``residualvit``: the gate regime on the ResidualViT-S shape, fixed per-image budgets).
Every rank takes its shard of one global batch; the averaged gradients must equal the single-process gradients of the whole
batch (computed on rank 0 with the same kernels), and one SGD step must leave identical parameters on every rank."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import weights as ow  # noqa: E402  (seeded synthetic weights only)
from peekvit_b200 import sharding  # noqa: E402
from peekvit_b200.finetune import FineTuner  # noqa: E402
from peekvit_b200.models import ResidualVisionTransformer, VisionTransformer  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gates = "residualvit" in sys.argv[1:]
    cfg = dict(image_size=224, patch_size=16, num_layers=12, num_heads=6, hidden_dim=384, mlp_dim=1536, num_classes=1000)
    B = 64 * world
    g = torch.Generator().manual_seed(3)
    images = torch.randn(B, 3, 224, 224, generator=g)
    labels = torch.randint(0, 1000, (B,), generator=g)
    budgets = torch.rand(B, generator=g) if gates else None
    if gates:
        cfg = dict(cfg, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
                   residual_layers=["attention+mlp"] * 12)
        sd = ow.calibrate_residual_gates(ow.make_state_dict("residualvit", cfg, seed=11), cfg, 0.5, images=images[:2])
    else:
        sd = ow.make_state_dict("vit", cfg, seed=11)
    kw = (lambda lo, hi: dict(budgets=budgets[lo:hi].to(dev))) if gates else (lambda lo, hi: {})

    def fresh():
        m = (ResidualVisionTransformer if gates else VisionTransformer)(**cfg)
        m.load_state_dict(sd)
        return m.to(dev).train()

    b, e = sharding.shard_range(B, rank, world)
    m = fresh()
    ft = FineTuner(m, micro_batch=32)
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=0.1)
    opt.zero_grad()
    loss, _ = ft.forward_backward(images[b:e].to(dev), labels[b:e].to(dev), **kw(b, e))
    grads = {n: p.grad.clone() for n, p in ft.params.items()}
    opt.step()
    res = {"world": world, "family": "residualvit (gate regime)" if gates else "vit", "trainable_tensors": len(ft.params), "local_loss": float(loss)}
    # reference: the whole batch in one process (every rank computes it; same kernels, no collective)
    m1 = fresh()
    ft1 = FineTuner(m1, micro_batch=32, process_group=None)
    import peekvit_b200.finetune as F
    saved = F.all_reduce_mean_
    F.all_reduce_mean_ = lambda params, group=None: 1
    try:
        ft1.forward_backward(images.to(dev), labels.to(dev), **kw(0, B))
    finally:
        F.all_reduce_mean_ = saved
    errs = {}
    for n, p in ft1.params.items():
        errs[n] = float((grads[n] - p.grad).abs().max() / p.grad.abs().max())
    res["grad_rel_err_vs_single_process"] = errs if len(errs) <= 8 else {"worst": max(errs.values()), "worst_name": max(errs, key=errs.get)}
    if world > 1:
        flat = torch.cat([p.detach().reshape(-1) for p in ft.params.values()])
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        res["params_identical_across_ranks_after_step"] = bool(torch.equal(lo, hi))
    ok = all(v < 2e-3 for v in errs.values()) and res.get("params_identical_across_ranks_after_step", True)
    res["ok"] = ok
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
