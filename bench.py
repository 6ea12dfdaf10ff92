#!/usr/bin/env python
"""bench.py — the headline benchmark: ViT-B/16 224px images/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-variants]

A "step" is one forward of the hot path over one synthetic batch: ``vit_b_16`` (configs[1]:
p16 D768 H12 F3072 L12, 1000 classes, random-init weights re-randomised so logits are not
identically zero), 2048 images per GPU, bf16 tensor-core operands with fp32 accumulation.
N > 1 is launched by torchrun, one rank per GPU; images shard by sample, weights are
replicated, the only collective is the all-reduce of the top-1 counts (``scaling: weak``).

Prints ONE JSON line on rank 0 (see README "Benchmark contract").  ``value`` is timed with the
inputs resident in HBM; ``e2e`` goes through the public module API with pinned host buffers,
H2D and D2H copies inside the timed region, over the same K steps.  Extra keys of the same line:

* ``strong_scaling``: BASELINE config B as written (2048 images TOTAL, 2048/N per GPU), device-resident and end to end;
* ``variants``: BASELINE configs A, C (four budgets), D (three budgets) and E, each with images/s, the ratio to the dense
  model of the same shape measured in the same run, the realised tokens / keep-fractions per layer and its own clock record;
* ``gpu_eager_reference``: the UNMODIFIED reference module (``baseline/_ref/peekvit``, vendored by
  ``baseline/vendor_reference.py``) in PyTorch eager on the same GPU, ``.cuda().eval().to(bfloat16)``;
* ``cpu_baseline``: the same unmodified reference module in fp32 on the host cores, on a bounded slice of the batch.

``--impl reference`` times that reference CPU forward alone (rank 0; other ranks exit).  Only when ``baseline/_ref`` is absent
do the CPU legs fall back to the oracle port (``kind: "port"``).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG_A = dict(image_size=224, patch_size=8, num_layers=4, num_heads=8, hidden_dim=256, mlp_dim=768, num_classes=10)
CFG_B = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
CFG_S = dict(image_size=224, patch_size=16, num_layers=12, num_heads=6, hidden_dim=384, mlp_dim=1536, num_classes=1000)
BATCH_PER_GPU = 2048
METRIC = "ViT-B/16 images/sec at 1/2/4/8 B200; tensor-pipe % of BF16 peak"


def gflop_per_image(cfg, tokens_per_layer=None, extra_tokens: int = 0) -> float:
    """2*P*Kp*D + sum_l(6nD^2 + 4n^2D + 2nD^2 + 4nDF) + 2DC (35.128 GFLOP for ViT-B/16); ``tokens_per_layer``
    = tokens entering each layer (survivor FLOPs of the budgeted variants)."""
    D, F, L, C = cfg["hidden_dim"], cfg["mlp_dim"], cfg["num_layers"], cfg["num_classes"]
    P = (cfg["image_size"] // cfg["patch_size"]) ** 2
    ns = tokens_per_layer if tokens_per_layer is not None else [P + 1 + extra_tokens] * L
    body = sum(6.0 * n * D * D + 4.0 * n * n * D + 2.0 * n * D * D + 4.0 * n * D * F for n in ns)
    return (2.0 * P * 3 * cfg["patch_size"] ** 2 * D + body + 2.0 * D * C) / 1e9


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=p.get("bf16_tflops_sustained", p.get("bf16_tflops")), hbm=p.get("hbm_gbs"), source="measured (sustained)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed regions (B200_PROFILING.md).  Every sample is stamped with the
    host time it arrived at, so one sampler serves the headline region and every variant (``window(t0, t1)``)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()

    def window(self, t0: float = 0.0, t1: float = float("inf")):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = [r for t, r in list(self.rows) if t0 <= t <= t1 + 0.06]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                if float(r[2]) < 300.0:        # idle samples (between regions) are not "under load"
                    continue
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            for r in rows:
                try:
                    sm.append(float(r[0])); mx = float(r[1])
                except Exception:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_weights_(model, seed: int = 4321) -> None:
    """the reference constructor zero-initialises head.*, the class token and every MHA bias, which makes
    random-init logits identically zero; the tensors it leaves at zero / one are re-drawn from a seeded generator (N(0, 0.02),
    LayerNorm gains 1 + N(0, 0.02)), everything else keeps the constructor's initialisation (under torch.manual_seed(0)).
    Gate projections of the budgeted models get a wider draw so that their scores are not degenerate."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if ".ln" in name and name.endswith("weight") or name.startswith("encoder.ln.weight"):
                p.copy_(1.0 + 0.02 * torch.randn(p.shape, generator=g))
            elif "residual_gate.projection.weight" in name or "budget_token_gate.weight" in name:
                p.copy_(torch.randn(p.shape, generator=g) / p.shape[-1] ** 0.5)
            elif "gating_network.gate.weight" in name:
                p.copy_(4.0 * (torch.rand(p.shape, generator=g) * 2 - 1) / p.shape[-1] ** 0.5)
            elif name.startswith("learnable_budget_token"):
                p.copy_(torch.randn(p.shape, generator=g))
            elif name.endswith("bias") or name in ("class_tokens", "class_token", "head.weight"):
                if name == "head.weight":
                    p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) / p.shape[1] ** 0.5)
                else:
                    p.copy_(0.02 * torch.randn(p.shape, generator=g))


def make_model(alias: str, cfg: dict, device):
    from peekvit_b200.models import build_model as pk_build
    torch.manual_seed(0)
    model = pk_build(alias, cfg)
    synthetic_weights_(model)
    return model.to(device).eval()


def build_model(device):
    """The measured arm builds its model and synthetic weights without touching ``oracle/``; the returned CPU state dict
    is what the reference legs load into the unmodified reference module."""
    from peekvit_b200.models import VisionTransformer
    torch.manual_seed(0)
    model = VisionTransformer(**CFG_B)
    synthetic_weights_(model)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    return model.to(device).eval(), sd


# ----------------------------------------------------------------------------------------------- reference legs
def load_reference_vit(sd=None):
    """The unmodified reference ``VisionTransformer`` (baseline/_ref/peekvit) with the benchmark's weights; None if the tree
    was not vendored (then the CPU legs fall back to the oracle port)."""
    try:
        from baseline.vendor_reference import import_reference
        cls = import_reference()["vit"]
    except Exception as e:      # noqa: BLE001
        return None, f"{type(e).__name__}: {e}"
    torch.manual_seed(0)
    ref = cls(**CFG_B)
    if sd is None:
        synthetic_weights_(ref)
    else:
        ref.load_state_dict(sd, strict=True)
    return ref.eval(), None


def host_threads() -> int:
    # torchrun exports OMP_NUM_THREADS=1 for its workers; the CPU legs use every host core this process may run on
    try:
        n = max(1, len(os.sched_getaffinity(0)))
    except Exception:
        n = max(1, os.cpu_count() or 1)
    torch.set_num_threads(n)
    return n


def cpu_forward_fn(sd):
    """(callable images -> logits, kind): the reference module on the host, else the oracle port."""
    ref, why = load_reference_vit(sd)
    if ref is not None:
        def f(x):
            with torch.no_grad():
                return ref(x)
        return f, "reference", None
    from oracle import peekvit_oracle as po, weights as ow
    sd = sd if sd is not None else ow.make_state_dict("vit", CFG_B, seed=4321)
    return (lambda x: po.forward("vit", sd, CFG_B, x)[0]), "port", why


def cpu_baseline(sd, budget_s=12.0, chunk=64, max_images=1024):
    """The reference forward (fp32, torch CPU ops, all host threads) on a bounded sample of the same workload."""
    cores = host_threads()
    f, kind, why = cpu_forward_fn(sd)
    images = torch.randn(chunk, 3, 224, 224, generator=torch.Generator().manual_seed(7))
    f(images[:8])                                              # warm the thread pool / allocator
    done, t0 = 0, time.perf_counter()
    while done < max_images and (time.perf_counter() - t0) < budget_s:
        f(images)
        done += chunk
    dt = time.perf_counter() - t0
    out = {"value": done / dt, "unit": "images/sec", "cores": cores, "kind": kind,
           "sample": f"{done} of {BATCH_PER_GPU} images of the ViT-B/16 224px batch, fp32, chunks of {chunk}, {dt:.1f} s, "
                     + ("unmodified reference VisionTransformer (baseline/_ref/peekvit)" if kind == "reference" else "oracle port")}
    if why:
        out["fallback_reason"] = why
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU forward of the same config on the host cores, rank 0 only."""
    if rank != 0:
        return
    cores = host_threads()
    torch.manual_seed(0)
    f, kind, why = cpu_forward_fn(None)
    chunk = 64
    images = torch.randn(chunk, 3, 224, 224, generator=torch.Generator().manual_seed(1234))
    for _ in range(max(args.warmup, 1)):
        f(images[:16])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        f(images)
    dt = time.perf_counter() - t0
    v = args.steps * chunk / dt
    sample = (f"{chunk} images per step x {args.steps} steps: a slice of the {BATCH_PER_GPU}-image batch, fp32, "
              + ("unmodified reference VisionTransformer.forward (baseline/_ref/peekvit)" if kind == "reference" else "oracle port"))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/sec", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "vit_b_16 224px forward (p16 D768 H12 F3072 L12 C1000), reference CPU implementation on the host cores",
                   "images_per_step": chunk, "full_batch": BATCH_PER_GPU},
        "cpu_baseline": {"value": v, "unit": "images/sec", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if why:
        line["cpu_baseline"]["fallback_reason"] = why
    print(json.dumps(line), flush=True)


def gpu_eager_reference(sd, dev, images, steps=5, chunk=256):
    """BASELINE.md §3 "the bar to beat": the unmodified reference module in PyTorch eager on this GPU, bf16 via
    ``.to(torch.bfloat16)``, ``torch.no_grad()``, inputs resident on the device, CUDA events."""
    ref, why = load_reference_vit(sd)
    if ref is None:
        return {"unavailable": why}
    try:
        ref = ref.to(dev).to(torch.bfloat16)
        x = images.to(torch.bfloat16)
        B = x.shape[0]

        def run():
            with torch.no_grad():
                for s in range(0, B, chunk):
                    ref(x[s:s + chunk])
        t_start = time.time()
        for _ in range(2):
            run()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        return {"value": B / ms * 1e3, "unit": "images/sec", "ms_per_step": ms, "steps": steps, "images_per_step": B, "chunk": chunk,
                "dtype": "bf16 (.to(torch.bfloat16))", "window": (t_start, time.time()),
                "api": "unmodified reference VisionTransformer (baseline/_ref/peekvit), torch eager (cuBLASLt / ATen)"}
    except Exception as e:      # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}
    finally:
        del ref
        torch.cuda.empty_cache()


# ----------------------------------------------------------------------------------------------- measured arm helpers
class Timer:
    """K steps bracketed by barrier + synchronize on both sides, CUDA events, max over ranks."""

    def __init__(self, dev, world):
        self.dev, self.world = dev, world
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def run(self, fn, steps: int, wall: bool = False):
        """-> (elapsed ms: max over ranks, host start time, host end time)."""
        self.barrier()
        t0 = time.time()
        p0 = time.perf_counter()
        self.e0.record()
        for _ in range(steps):
            fn()
        self.e1.record()
        self.barrier()
        ms = self.e0.elapsed_time(self.e1)
        if wall:
            ms = max(ms, (time.perf_counter() - p0) * 1e3)
        t1 = time.time()
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t0, t1


def calibrate_residual_gates_(model, budget: float, probe: torch.Tensor, target: float, iters: int = 12) -> None:
    """with random gates the realised keep-fraction is 0 or 1 whatever the budget.  Layer by layer, the
    gate's projection bias is bisected ON THE DEVICE (the product path's own forward and published ``block.mask``) until the
    probe batch keeps ~``target`` of its image tokens at ``budget``.  The bias is written into the live parameter."""
    from peekvit_b200 import runner
    model.set_budget(budget)
    graphs, runner.USE_CUDA_GRAPHS = runner.USE_CUDA_GRAPHS, False
    try:
        for i, blk in enumerate(model.encoder.layers):
            if getattr(blk, "skip", None) != "attention+mlp":
                continue
            lw = runner.packed(model).layers[i]
            lo, hi = -16.0, 16.0
            for _ in range(iters):
                mid = 0.5 * (lo + hi)
                lw.extra["gate_b"] = mid              # the packed copy of projection.bias (engine.pack_model)
                model(probe)
                frac = float((blk.mask > 0).float().mean())
                if frac > target:
                    hi = mid
                else:
                    lo = mid
            with torch.no_grad():
                blk.residual_gate.projection.bias.fill_(0.5 * (lo + hi))
    finally:
        runner.USE_CUDA_GRAPHS = graphs


def calibrate_avit_center_(model, probe: torch.Tensor, target_layers: float, iters: int = 12) -> float:
    """avit_s_16_224.yaml ships gate_scale 10 / gate_center 30 for DeiT-S weights; with random-init weights no token would ever
    halt.  The halting centre is bisected on the device until a probe batch's tokens run ~``target_layers`` layers on average."""
    from peekvit_b200 import runner
    graphs, runner.USE_CUDA_GRAPHS = runner.USE_CUDA_GRAPHS, False
    model.pk_early_exit = False                 # full per-token bookkeeping while calibrating
    try:
        lo, hi = -40.0, 40.0
        for _ in range(iters):
            mid = 0.5 * (lo + hi)
            for blk in model.encoder.layers:
                blk.gate_center = mid
            model(probe)
            mean_layers = float(model.encoder.counter_token.float().mean())
            if mean_layers > target_layers:         # a lower centre halts earlier
                hi = mid
            else:
                lo = mid
        for blk in model.encoder.layers:
            blk.gate_center = 0.5 * (lo + hi)
    finally:
        runner.USE_CUDA_GRAPHS = graphs
        model.pk_early_exit = True
    return 0.5 * (lo + hi)


def run_variants(dev, rank, world, timer: Timer, sampler, images_full, steps_hint: int, min_seconds: float = 0.8):
    """BASELINE configs A, C, D, E next to their dense baselines: device-resident synthetic images, CUDA events, max over ranks.
    Every entry is timed for at least ``min_seconds`` so that its own clock record holds several samples."""
    from peekvit_b200 import ops, runner
    out = {}
    B = images_full.shape[0]

    def timed(model, images, name, extra=None, keep_aux=False):
        for _ in range(3):
            model(images)
        ms1, _, _ = timer.run(lambda: model(images), 2)
        steps = max(steps_hint, 3, int(math.ceil(min_seconds * 1e3 / max(ms1 / 2, 1e-3))))
        steps = min(steps, 400)
        ms, t0, t1 = timer.run(lambda: model(images), steps)
        v = world * images.shape[0] * steps / (ms * 1e-3)
        entry = {"value": v, "unit": "images/sec", "images_per_gpu_per_step": images.shape[0], "steps": steps, "ms_per_step": ms / steps,
                 "micro_batch": min(runner._micro_batch(model, images.shape[0]), images.shape[0])}
        if rank == 0 and sampler is not None:
            entry["clocks"] = sampler.window(t0, t1)
        entry.update(extra or {})
        out[name] = entry
        return v

    # ---- dense baselines
    vitb = make_model("vit", CFG_B, dev)
    dense_b = timed(vitb, images_full, "vit_b_16_dense", {"config": "B", "gflop_per_image": gflop_per_image(CFG_B)})
    del vitb
    vits = make_model("vit", CFG_S, dev)
    vits.pk_micro_batch = runner.SPARSE_MICRO_BATCH        # the denominator of configs C / E runs at the micro-batch its numerators use (+2 % for the dense model)
    dense_s = timed(vits, images_full, "vit_s_16_dense", {"gflop_per_image": gflop_per_image(CFG_S)})
    del vits

    # ---- config A: vit_tiny p8 (785 tokens, head_dim 32), batch 64, the reference's CPU-runnable fp32 case
    vita = make_model("vit", CFG_A, dev)
    xa = images_full[:64]
    timed(vita, xa, "A_vit_tiny_p8_b64_bf16", {"config": "A", "gflop_per_image": gflop_per_image(CFG_A)})
    vita.pk_precision = "fp32"
    timed(vita, xa, "A_vit_tiny_p8_b64_fp32_mode", {"config": "A", "gflop_per_image": gflop_per_image(CFG_A),
                                                    "note": "pk_precision='fp32' (the dtype config A is stated in)"})
    del vita

    # ---- config D: RankViT on the ViT-B shape, rank layers [3, 6, 9]
    cfg = dict(CFG_B, rankvit_layers=[3, 6, 9])
    m = make_model("RankVisionTransformer", cfg, dev)
    for budget in (0.5, 0.4, 0.25):
        m.set_budget(budget)
        aux = {}
        runner.run(m, images_full[:32], aux)
        lens = [int(n) for n in aux.get("seq_lens", [])]
        v = timed(m, images_full, f"D_rankvit_b_budget{budget}", None)
        out[f"D_rankvit_b_budget{budget}"].update(config="D", budget=budget, x_dense=v / dense_b, tokens_per_layer=lens,
                                                  gflop_per_image=gflop_per_image(CFG_B, lens))
    del m

    # ---- config C: ResidualViT, ViT-S shape (residualdeit_s_16_224.yaml kwargs), learnable budget token, every layer gated
    cfg = dict(CFG_S, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
               add_input=False, residual_layers=["attention+mlp"] * 12)
    for budget in (0.2, 0.4, 0.8, 1.0):
        m = make_model("residualvit", cfg, dev)
        calibrate_residual_gates_(m, budget, images_full[:32], target=min(budget, 0.97))
        m.set_budget(budget)
        v = timed(m, images_full, f"C_residualvit_s_budget{budget}", None)
        keep = [round(float((blk.mask > 0).float().mean()), 3) for blk in m.encoder.layers if getattr(blk, "mask", None) is not None]
        toks = [2 + 196 * k for k in keep]
        out[f"C_residualvit_s_budget{budget}"].update(
            config="C", budget=budget, x_dense=v / dense_s, keep_fraction_per_layer=keep, mean_keep_fraction=round(sum(keep) / max(len(keep), 1), 3),
            gflop_per_image=gflop_per_image(CFG_S, toks), gates="projection bias bisected per layer on the device so the keep-fraction ~ budget")
        del m

    # ---- config E: A-ViT halting and MoE expert MLPs on the ViT-S shape
    cfg = dict(CFG_S, eps=0.01, gate_scale=10, gate_center=30)          # avit_s_16_224.yaml; the centre is re-calibrated below
    m = make_model("adavit", cfg, dev)
    center = calibrate_avit_center_(m, images_full[:32], target_layers=7.0)
    v = timed(m, images_full, "E_avit_s", None)
    m.pk_early_exit = False
    m(images_full[:64])
    cnt = m.encoder.counter_token
    out["E_avit_s"].update(config="E", x_dense=v / dense_s, gate_scale=10, gate_center=round(center, 3),
                           mean_layers_per_token=float(cnt.float().mean()) if cnt is not None else None)
    del m
    cfg = dict(CFG_S, mlp_moes=[4] * 12)
    m = make_model("vitmoe", cfg, dev)
    v = timed(m, images_full, "E_moevit_s_4experts", None)
    out["E_moevit_s_4experts"].update(config="E", x_dense=v / dense_s, experts=4,
                                      note="same FLOPs as the dense ViT-S (arg-max routing); the reference evaluates all experts densely")
    del m
    # ---- fine-tuning step: forward + cross-entropy + backward of the class-token / head regime on ViT-B/16
    # (train/train.py:97-127 with the backbone frozen), gradients averaged over the ranks by one NCCL all-reduce per step
    try:
        from peekvit_b200.finetune import FineTuner
        m = make_model("vit", CFG_B, dev)
        m.train()
        ft = FineTuner(m, micro_batch=256)
        nb = 512
        xb, yb = images_full[:nb], torch.randint(0, CFG_B["num_classes"], (nb,), device=dev)
        opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=1e-3)

        def train_step():
            opt.zero_grad()
            ft.forward_backward(xb, yb)
            opt.step()
        for _ in range(2):
            train_step()
        ms, t0, t1 = timer.run(train_step, 3)
        entry = {"value": world * nb * 3 / (ms * 1e-3), "unit": "images/sec", "images_per_gpu_per_step": nb, "steps": 3,
                 "ms_per_step": ms / 3, "config": "f4", "regime": "class_tokens + head trainable, backbone frozen (activation gradients only)",
                 "includes": "forward, cross-entropy, backward, gradient all-reduce, SGD step"}
        if rank == 0 and sampler is not None:
            entry["clocks"] = sampler.window(t0, t1)
        out["F_finetune_vit_b_16_cls_head"] = entry
        del m, ft, opt
    except Exception as e:      # noqa: BLE001  (a bench line must not die on the optional leg)
        out["F_finetune_vit_b_16_cls_head"] = {"error": str(e)[:300]}
    # ---- the gate regime on the ResidualViT-S shape (config C kwargs): training-mode forward with one sampled budget per image,
    # soft masks on all tokens, backward through the masks into the gate projections / budget-token gates / learnable budget
    # token / class token / head; cross-entropy only (a mask regulariser is the caller's torch function)
    try:
        from peekvit_b200.finetune import FineTuner
        cfg = dict(CFG_S, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
                   add_input=False, residual_layers=["attention+mlp"] * 12)
        m = make_model("residualvit", cfg, dev)
        calibrate_residual_gates_(m, 0.5, images_full[:32], target=0.5)
        m.train()
        ft = FineTuner(m, micro_batch=512)
        nb = 512
        xb, yb = images_full[:nb], torch.randint(0, cfg["num_classes"], (nb,), device=dev)
        opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=1e-3)

        def gate_step():
            opt.zero_grad()
            ft.forward_backward(xb, yb)
            opt.step()
        for _ in range(2):
            gate_step()
        ms, t0, t1 = timer.run(gate_step, 3)
        entry = {"value": world * nb * 3 / (ms * 1e-3), "unit": "images/sec", "images_per_gpu_per_step": nb, "steps": 3,
                 "ms_per_step": ms / 3, "config": "f4", "trainable_tensors": len(ft.params),
                 "regime": "gates + budget-token gates + learnable budget token + class token + head trainable, backbone frozen",
                 "includes": "budget sampling, training-mode forward (all tokens, soft masks), cross-entropy, backward, gradient all-reduce, SGD step"}
        if rank == 0 and sampler is not None:
            entry["clocks"] = sampler.window(t0, t1)
        out["F_finetune_residualvit_s_gates"] = entry
        del m, ft, opt
    except Exception as e:      # noqa: BLE001
        out["F_finetune_residualvit_s_gates"] = {"error": str(e)[:300]}
    out["device_flag"] = ops.device_flag()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="images per GPU per step")
    ap.add_argument("--micro-batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: peekvit_b200 has no CPU path (use --impl reference for the CPU arm)")
    import torch.distributed as dist
    from peekvit_b200 import ops, runner
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    model, sd = build_model(dev)
    if args.micro_batch:
        model.pk_micro_batch = args.micro_batch
    B = args.batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    images = torch.randn(B, 3, 224, 224, device=dev, generator=g)              # 1.2 GB > 126 MB L2
    labels = torch.randint(0, CFG_B["num_classes"], (B,), device=dev, generator=g)
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    local_counts = torch.zeros(2, dtype=torch.int64, device=dev)
    timer = Timer(dev, world)

    def make_step(x, y):
        def step():
            logits = model(x)
            # the eval loop's accuracy count (validate/test.py:120-127) fused on the device (pk_argmax_count); the only
            # cross-GPU exchange is the all-reduce of the two counters
            local_counts.zero_()
            ops.argmax_count(logits, y, local_counts)
            if world > 1:
                counts.copy_(local_counts)
                dist.all_reduce(counts, op=dist.ReduceOp.SUM)
            return logits
        return step

    step = make_step(images, labels)

    # pinned host copies of the batch for the end-to-end leg (allocated and warmed up here so that the timed regions
    # below run back to back in the same thermal / power state)
    host_images = torch.empty(B, 3, 224, 224, dtype=torch.float32, pin_memory=True)
    host_images.copy_(images)
    host_logits = torch.empty(B, CFG_B["num_classes"], dtype=torch.float32, pin_memory=True)
    # a prefetching loader holds the next batch in a second pinned buffer while the current one runs: the e2e loop alternates
    # between two host buffers and names the next one, so every step still copies all of its inputs -- the first chunk under
    # the previous step's last micro-batch instead of in front of its own
    host_images2 = torch.empty_like(host_images, pin_memory=True)
    host_images2.copy_(host_images)
    host_pair = [host_images, host_images2]
    e2e_count = [0]

    def e2e_step(pair=host_pair, logits=host_logits, count=e2e_count):
        i = count[0] & 1
        count[0] += 1
        model.forward_host(pair[i], logits, next_host=pair[i ^ 1])
        if world > 1:
            # the same per-step exchange as the device-resident step (the accuracy counters): both legs run the ranks in
            # lockstep, so the end-to-end number cannot come out ahead by skipping the collective
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    for _ in range(2):
        e2e_step()
    # the clock sampler starts BEFORE the warm-up (idle samples are filtered by power draw), so that the timed region follows
    # the warm-up steps without an idle gap: a pause right before it lets the GPU boost above its sustained clocks
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.start()
        time.sleep(0.3)
    timer.barrier()
    for _ in range(args.warmup):
        step()
    ops.launch_count = 0
    elapsed_ms, t_head0, _ = timer.run(step, args.steps)
    launches = ops.launch_count

    # ---- e2e: pinned host batch -> module API -> host logits, copies inside the timed region, same number of steps
    e2e_steps = args.steps
    e2e_ms, _, t_head1 = timer.run(e2e_step, e2e_steps, wall=True)

    # ---- roofline pass: the same K steps again with CUDA events around every GEMM launch (the per-launch
    # events cannot be recorded from inside the CUDA-graph replay the timed region uses, so this pass runs
    # the identical launch sequence eagerly; clocks are sampled over both regions)
    ops.gemm_timeline = []
    eager_ms, _, t_head2 = timer.run(step, args.steps)
    timeline, ops.gemm_timeline = ops.gemm_timeline, None
    clocks = sampler.window(t_head0, t_head2) if sampler is not None else None
    flag = ops.device_flag()

    # ---- strong scaling: BASELINE config B as written, 2048 images in TOTAL -> 2048 / N per GPU (identical to the weak
    # numbers at N = 1)
    strong = None
    if world > 1:
        Bs = max(1, BATCH_PER_GPU // world)
        xs, ys = images[:Bs], labels[:Bs]
        hs, hl = host_images[:Bs], host_logits[:Bs]
        s_pair, s_count = [host_images[:Bs], host_images2[:Bs]], [0]
        sstep = make_step(xs, ys)
        for _ in range(3):
            sstep()
            e2e_step(s_pair, hl, s_count)
        s_steps = max(args.steps, 20)
        s_ms, st0, _ = timer.run(sstep, s_steps)
        s_e2e_ms, _, st1 = timer.run(lambda: e2e_step(s_pair, hl, s_count), s_steps, wall=True)
        strong = {"scaling": "strong", "global_batch": Bs * world, "images_per_gpu_per_step": Bs, "steps": s_steps,
                  "value": world * Bs * s_steps / (s_ms * 1e-3), "unit": "images/sec", "ms_per_step": s_ms / s_steps,
                  "e2e": {"value": world * Bs * s_steps / (s_e2e_ms * 1e-3), "unit": "images/sec",
                          "h2d_bytes_per_step": hs.numel() * 4, "d2h_bytes_per_step": hl.numel() * 4},
                  "clocks": sampler.window(st0, st1) if sampler is not None else None,
                  "note": "BASELINE config B (2048 images total); efficiency = value / (N x the N=1 value of the weak line)"}

    # ---- same end-to-end call fed with uint8 HWC images (ToTensor + Normalize fused into the im2col):
    # informational, the headline e2e above is the reference-facing float API
    host_u8 = torch.empty(B, 224, 224, 3, dtype=torch.uint8, pin_memory=True)
    host_u8.random_(0, 256)
    for _ in range(2):
        model.forward_host(host_u8, host_logits)
    u8_steps = max(2, min(args.steps, 5))
    e2e_u8_ms, _, _ = timer.run(lambda: model.forward_host(host_u8, host_logits), u8_steps)
    del host_u8

    # ---- the other arithmetic modes of the same model: informational, every rank runs them (keeps the ranks in step)
    modes = {}
    for prec, nimg in (("fp32", 256), ("bf16x2", 1024)):
        if prec not in runner.PRECISIONS:
            continue
        model.pk_precision = prec
        sl = images[:nimg]
        for _ in range(2):
            hi = model(sl)
        ms, _, _ = timer.run(lambda: model(sl), 3)
        model.pk_precision = "bf16"
        quick = model(sl)
        modes[prec] = {"value": world * nimg * 3 / (ms * 1e-3), "unit": "images/sec", "images_per_gpu_per_step": nimg,
                       "bf16_vs_this_mode_rel_diff": float(((quick - hi).abs().max() / hi.abs().max()).item())}
    model.pk_precision = "bf16"

    # ---- BASELINE configs A, C, D, E (driver-observed, each entry with its own clock record)
    variants = None
    if not args.no_variants:
        variants = run_variants(dev, rank, world, timer, sampler, images, steps_hint=min(args.steps, 10))

    # ---- the unmodified reference module, torch eager, same GPU (rank 0; the other ranks wait at the barrier)
    gpu_ref = None
    if rank == 0 and not args.no_gpu_reference:
        gpu_ref = gpu_eager_reference(sd, dev, images)
        if sampler is not None and "window" in gpu_ref:
            w = gpu_ref.pop("window")
            gpu_ref["clocks"] = sampler.window(*w)
    timer.barrier()
    if sampler is not None:
        sampler.stop()

    if rank == 0:
        peaks = measured_peaks()
        gemm_ms = sum(a.elapsed_time(b) for a, b, _ in timeline)
        gemm_flops = sum(f for _, _, f in timeline)
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("gemm_dram_bytes_per_launch")
        value = world * B * args.steps / (elapsed_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "vit_b_16 224px forward (p16 D768 H12 F3072 L12 C1000), random-init weights re-randomised (seed 4321)",
                       "images_per_gpu_per_step": B, "global_batch": world * B, "micro_batch": int(getattr(model, "pk_micro_batch", runner.DEFAULT_MICRO_BATCH)),
                       "parallelism": f"dp{world} (sample-sharded, replicated weights)",
                       "l2": "inputs 1.2 GB/step and activations per micro-batch exceed the 126 MB L2",
                       "launch": "CUDA graph replay per micro-batch" if runner.USE_CUDA_GRAPHS else "eager launches",
                       "accumulate": "fp32 (TMEM), fp32 residual stream / LayerNorm / softmax statistics"},
            "model_tflops": value * gflop_per_image(CFG_B) / 1e3,
            "model_frac_of_peak": value / world * gflop_per_image(CFG_B) / 1e3 / peaks["tflops"],
            "gpu_launches": launches,
            "clocks": clocks,
            "device_flag": flag,
            "e2e": {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": "images/sec", "steps": e2e_steps,
                    "h2d_bytes_per_step": host_images.numel() * 4, "d2h_bytes_per_step": host_logits.numel() * 4,
                    "api": "VisionTransformer.forward_host(pinned images, next_host=the loader's next pinned batch) -> pinned logits; two host buffers alternate",
                    "timing": "host wall clock around K calls, each ending in a stream synchronise after the D2H copy of its logits",
                    "note": "every H2D chunk overlaps compute (the next batch's first chunk is staged under the previous batch's "
                            "last micro-batch), so this sits within run-to-run clock noise (sw_power_cap, ~1 %) of `value`, "
                            "which is timed separately with CUDA events and additionally runs the accuracy count"},
            "strong_scaling": strong,
            "precision_modes": modes,
            "e2e_uint8_input": {"value": world * B * u8_steps / (e2e_u8_ms * 1e-3), "unit": "images/sec",
                                "h2d_bytes_per_step": B * 224 * 224 * 3, "note": "same call with uint8 HWC images; ToTensor + Normalize fused into the im2col"},
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_pair_kernel (tcgen05 cta_group::2: patch embedding / QKV / out-proj / fc1+GELU / fc2)",
                         "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                         "peak_source": peaks["source"], "traffic": traffic, "launches": len(timeline),
                         "avg_launch_ms": gemm_ms / max(len(timeline), 1), "share_of_step": gemm_ms / eager_ms,
                         "timing": "CUDA events around every GEMM launch in a second, eager pass over the same K steps "
                                   f"({eager_ms / args.steps:.2f} ms/step; the timed region replays CUDA graphs)"},
            "variants": variants,
            "gpu_eager_reference": gpu_ref,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(sd)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
