"""Ragged tcgen05 attention (impl 3) against the general mma.sync kernel (impl 1) and, on uniform shapes, the dense tcgen05
kernel (impl 2): python tools/attn_ragged_bench.py [--json out.json]

Shapes: one ResidualViT-S layer at budget 0.4 (512 images x 6 heads, ~80 live rows, multiplicities + virtual key), one A-ViT-S
layer, the pruned RankViT-B layers (uniform 99 / 50 / 26 / 14 tokens, 12 heads), the dense ViT-B / ViT-S layer.
FLOPs = 4 * len * keys * 64 per (sample, head).  CUDA events, 20 launches back to back after 3 warm-ups."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops  # noqa: E402

DEV = "cuda:0"
CAP = int(os.environ.get("ATT_BENCH_CAP", "199"))


def timed(fn, n=20):
    """us per launch, n launches replayed from one CUDA graph (eager launches of a 40 us kernel measure the host's
    ctypes + tensor-map encoding time instead of the kernel)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3          # us


def case(name, lens, H, mult, extra, res):
    dh, D = 64, H * 64
    B, rows = len(lens), sum(lens)
    g = torch.Generator(device=DEV).manual_seed(1)
    qkv = (torch.randn(rows, 3 * D, device=DEV, generator=g)).to(torch.bfloat16)
    out = torch.zeros(rows, D, device=DEV, dtype=torch.bfloat16)
    uniform = len(set(lens)) == 1 and not mult and not extra
    # the model launches with the static upper bound of the row count per sample (199 / 197), not the realised maximum
    kw = dict(seq_len=lens[0]) if uniform else dict(cu_seqlens=torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), device=DEV, dtype=torch.int32),
                                                    max_seq_len=max(max(lens), CAP))
    if mult:
        km = torch.ones(rows, device=DEV)
        km[torch.tensor(lens).cumsum(0).to(DEV) - 1] = 37.0          # the ghost row of every sample
        kw["key_mult"] = km
    if extra:
        kw["extra_kv"] = (torch.randn(2 * D, device=DEV, generator=g) * 0.3).to(torch.bfloat16)
        kw["extra_mult"] = torch.full((B,), 100.0, device=DEV)
    flops = sum(4.0 * n * (n + (1 if extra else 0)) * 64 * H for n in lens)
    entry = {"samples": B, "heads": H, "mean_len": rows / B, "gflop": flops / 1e9}
    for impl, label in ((3, "tcgen05_ragged"), (1, "mma_sync_general")) + (((2, "tcgen05_dense"),) if uniform and 17 <= lens[0] <= 256 else ()):
        try:
            us = timed(lambda: ops.attention(qkv, out, B, H, dh, impl=impl, **kw))
            entry[label] = {"us": us, "tflops": flops / us / 1e6}
        except Exception as e:      # noqa: BLE001
            entry[label] = {"error": str(e)[:200]}
    if not uniform and max(lens) + (1 if extra else 0) <= 128:
        kw4 = dict(kw, max_seq_len=max(lens))
        try:
            us = timed(lambda: ops.attention(qkv, out, B, H, dh, impl=4, **kw4))
            entry["tcgen05_quad"] = {"us": us, "tflops": flops / us / 1e6}
        except Exception as e:      # noqa: BLE001
            entry["tcgen05_quad"] = {"error": str(e)[:200]}
    entry["device_flag"] = ops.device_flag()
    res[name] = entry
    print(name, json.dumps(entry), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    res = {}
    g = torch.Generator().manual_seed(0)
    lens04 = (torch.randint(40, 125, (512,), generator=g)).tolist()
    case("residualvit_s_b0.4_layer", lens04, 6, True, True, res)
    lens02 = (torch.randint(10, 80, (512,), generator=g)).tolist()
    case("residualvit_s_b0.2_layer", lens02, 6, True, True, res)
    lens08 = (torch.randint(130, 199, (512,), generator=g)).tolist()
    case("residualvit_s_b0.8_layer", lens08, 6, True, True, res)
    case("avit_s_layer", (torch.randint(30, 198, (512,), generator=g)).tolist(), 6, False, True, res)
    for n in (99, 80, 50, 33, 26, 14):
        case(f"rankvit_b_uniform_{n}", [n] * 512, 12, False, False, res)
    for n in (50, 26):
        case(f"ragged_layout_uniform_{n}", [n] * 511 + [n - 1], 12, False, False, res)
    case("vit_b_uniform_197", [197] * 512, 12, False, False, res)
    case("vit_s_uniform_198", [198] * 512, 6, False, False, res)
    if args.json:
        os.makedirs(os.path.dirname(args.json), exist_ok=True)
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
