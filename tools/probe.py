"""On-GPU kernel probe: runs one kernel family against a torch fp32 expression and prints error
statistics (with enough structure to diagnose a wrong descriptor or layout remotely).

    python tools/probe.py gemm|ln|attn|rows|rank|model [--json gpurun_out/probe_x.json]

Each family is meant to run in its own process under `timeout` (tools/run_probes.sh).
"""
from __future__ import annotations

import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from peekvit_b200 import ops  # noqa: E402
from peekvit_b200._lib import (PK_EPI_BIAS_BF16, PK_EPI_BIAS_F32, PK_EPI_BIAS_GELU_BF16, PK_EPI_BIAS_RESID_F32)  # noqa: E402

DEV = "cuda:0"
RESULTS = {}


def rel_err(got, ref):
    got, ref = got.float(), ref.float()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def report(name, **kw):
    RESULTS[name] = kw
    print(name, json.dumps(kw), flush=True)


def time_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def probe_gemm():
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    shapes = [(128, 192, 64), (128, 128, 128), (256, 256, 64), (300, 768, 768), (1000, 2304, 768), (1576, 3072, 768),
              (257, 768, 3072), (64, 128, 192), (197 * 64, 2304, 768)]
    for (M, N, K) in shapes:
        a = (torch.randn(M, K, device=DEV) * 0.5).to(torch.bfloat16)
        w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn(N, device=DEV) * 0.1
        ref = a.float() @ w.float().t() + bias
        for bn in (0, 128, 192, 256):
            out = torch.full((M, N), float("nan"), device=DEV)
            ops.gemm(a, w, bias, out, PK_EPI_BIAS_F32, block_n=bn)
            flag = ops.device_flag()
            err = rel_err(out, ref)
            bad = ~((out - ref).abs() <= 1e-3 * ref.abs().max())
            extra = {}
            if bad.any():
                rows = bad.any(1).nonzero().flatten()
                cols = bad.any(0).nonzero().flatten()
                extra = dict(bad_frac=bad.float().mean().item(), bad_rows=rows[:8].tolist(), n_bad_rows=int(rows.numel()),
                             bad_cols=cols[:8].tolist(), n_bad_cols=int(cols.numel()), nan=int(torch.isnan(out).sum()),
                             sample_got=out[rows[0], cols[0]].item(), sample_ref=ref[rows[0], cols[0]].item())
            report(f"gemm_f32_M{M}_N{N}_K{K}_bn{bn}", err=err, flag=flag, **extra)
            if flag:
                return
    # epilogues on one shape
    M, N, K = 1000, 768, 768
    a = (torch.randn(M, K, device=DEV) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV) * 0.1
    acc = a.float() @ w.float().t() + bias
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, bias, out, PK_EPI_BIAS_BF16)
    report("gemm_epi_bf16", err=rel_err(out, acc), flag=ops.device_flag())
    ops.gemm(a, w, bias, out, PK_EPI_BIAS_GELU_BF16)
    report("gemm_epi_gelu", err=rel_err(out, torch.nn.functional.gelu(acc)), flag=ops.device_flag())
    x = torch.randn(M, N, device=DEV)
    rs = torch.rand(M, device=DEV)
    x0 = x.clone()
    ops.gemm(a, w, bias, x, PK_EPI_BIAS_RESID_F32, resid=x, rowscale=rs)
    report("gemm_epi_resid_rowscale_inplace", err=rel_err(x, rs[:, None] * acc + x0), flag=ops.device_flag())
    # patch-embed row remap: 3 groups of 10 rows -> seq 13, offset 2, + pos table
    G, P, seq, off = 5, 60, 66, 3
    a = (torch.randn(G * P, 192, device=DEV)).to(torch.bfloat16)
    w = (torch.randn(256, 192, device=DEV) / 14).to(torch.bfloat16)
    bias = torch.randn(256, device=DEV) * 0.1
    pos = torch.randn(seq, 256, device=DEV)
    x = torch.zeros(G * seq, 256, device=DEV)
    ops.gemm(a, w, bias, x, PK_EPI_BIAS_RESID_F32, resid=pos, rows_per_group=P, group_stride=seq, group_offset=off, resid_is_pos=True)
    ref = torch.zeros(G, seq, 256, device=DEV)
    ref[:, off:off + P] = (a.float() @ w.float().t() + bias).view(G, P, 256) + pos[off:off + P]
    report("gemm_patch_remap", err=rel_err(x, ref.view(G * seq, 256)), flag=ops.device_flag())
    # device-side M
    M, N, K = 1000, 384, 384
    a = (torch.randn(M, K, device=DEV) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    out = torch.zeros(M, N, device=DEV)
    mdev = torch.tensor([333], device=DEV, dtype=torch.int32)
    ops.gemm(a, w, None, out, PK_EPI_BIAS_F32, m_dev=mdev)
    ref = a.float() @ w.float().t()
    report("gemm_mdev", err=rel_err(out[:333], ref[:333]), untouched=bool((out[333:] == 0).all()), flag=ops.device_flag())
    # timing on the ViT-B shapes (B=256 images)
    Mbig = 197 * 256
    for (N, K, epi, name) in [(2304, 768, PK_EPI_BIAS_BF16, "qkv"), (768, 768, PK_EPI_BIAS_RESID_F32, "proj"),
                              (3072, 768, PK_EPI_BIAS_GELU_BF16, "fc1"), (768, 3072, PK_EPI_BIAS_RESID_F32, "fc2")]:
        a = (torch.randn(Mbig, K, device=DEV) * 0.5).to(torch.bfloat16)
        w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn(N, device=DEV) * 0.1
        bf = epi in (PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16)
        out = torch.empty(Mbig, N, device=DEV, dtype=torch.bfloat16 if bf else torch.float32)
        resid = None if bf else out
        for bn in (128, 192, 256):
            ms = time_ms(lambda: ops.gemm(a, w, bias, out, epi, resid=resid, block_n=bn))
            report(f"gemm_time_{name}_bn{bn}", ms=ms, tflops=2.0 * Mbig * N * K / ms / 1e9, flag=ops.device_flag())
        ms = time_ms(lambda: torch.nn.functional.linear(a, w))
        report(f"cublas_time_{name}", ms=ms, tflops=2.0 * Mbig * N * K / ms / 1e9)


def probe_gemm2():
    """CTA-pair (cta_group::2) kernel: every epilogue, partial tiles, device-side M, then timing vs the
    single-CTA kernel and cuBLAS on the ViT-B shapes."""
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    shapes = [(256, 256, 64), (512, 256, 128), (300, 768, 768), (1000, 2304, 768), (1576, 3072, 768), (257, 768, 3072),
              (1000, 384, 384), (777, 1152, 384), (197 * 64, 2304, 768), (600, 1000, 768)]
    for (M, N, K) in shapes:
        a = (torch.randn(M, K, device=DEV) * 0.5).to(torch.bfloat16)
        w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn(N, device=DEV) * 0.1
        acc = a.float() @ w.float().t() + bias
        for bn in (0, 128, 192, 256):
            out = torch.full((M, N), float("nan"), device=DEV)
            ops.gemm(a, w, bias, out, PK_EPI_BIAS_F32, block_n=bn, cta_pair=2)
            flag = ops.device_flag()
            err = rel_err(out, acc)
            bad = ~((out - acc).abs() <= 1e-3 * acc.abs().max())
            extra = {}
            if bad.any():
                rows = bad.any(1).nonzero().flatten()
                cols = bad.any(0).nonzero().flatten()
                extra = dict(bad_frac=bad.float().mean().item(), bad_rows=rows[:8].tolist(), n_bad_rows=int(rows.numel()),
                             bad_cols=cols[:8].tolist(), n_bad_cols=int(cols.numel()), nan=int(torch.isnan(out).sum()),
                             sample_got=out[rows[0], cols[0]].item(), sample_ref=acc[rows[0], cols[0]].item())
            report(f"pair_f32_M{M}_N{N}_K{K}_bn{bn}", err=err, flag=flag, **extra)
            if flag:
                return
        for bn in (0, 128, 192):
            outb = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
            ops.gemm(a, w, bias, outb, PK_EPI_BIAS_BF16, block_n=bn, cta_pair=2)
            e1 = rel_err(outb, acc)
            ops.gemm(a, w, bias, outb, PK_EPI_BIAS_GELU_BF16, block_n=bn, cta_pair=2)
            e2 = rel_err(outb, torch.nn.functional.gelu(acc))
            x = torch.randn(M, N, device=DEV)
            rs = torch.rand(M, device=DEV)
            x0 = x.clone()
            ops.gemm(a, w, bias, x, PK_EPI_BIAS_RESID_F32, resid=x, rowscale=rs, block_n=bn, cta_pair=2)
            e3 = rel_err(x, rs[:, None] * acc + x0)
            report(f"pair_epi_M{M}_N{N}_K{K}_bn{bn}", bf16=e1, gelu=e2, resid=e3, flag=ops.device_flag())
    # device-side M: rows >= m_dev stay untouched
    M, N, K = 1000, 768, 384
    a = (torch.randn(M, K, device=DEV) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    ref = a.float() @ w.float().t()
    for md in (0, 1, 333, 512, 1000):
        mdev = torch.tensor([md], device=DEV, dtype=torch.int32)
        out = torch.zeros(M, N, device=DEV)
        ops.gemm(a, w, None, out, PK_EPI_BIAS_F32, m_dev=mdev, cta_pair=2)
        outb = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
        ops.gemm(a, w, None, outb, PK_EPI_BIAS_BF16, m_dev=mdev, cta_pair=2)
        x = torch.ones(M, N, device=DEV)
        ops.gemm(a, w, None, x, PK_EPI_BIAS_RESID_F32, resid=x, m_dev=mdev, cta_pair=2)
        report(f"pair_mdev_{md}", err=rel_err(out[:md], ref[:md]) if md else 0.0, errb=rel_err(outb[:md], ref[:md]) if md else 0.0,
               errr=rel_err(x[:md], ref[:md] + 1) if md else 0.0,
               untouched=bool((out[md:] == 0).all() and (outb[md:] == 0).all() and (x[md:] == 1).all()), flag=ops.device_flag())
    # timing on the ViT-B shapes (B=256 images)
    Mbig = 197 * 256
    for (N, K, epi, name) in [(2304, 768, PK_EPI_BIAS_BF16, "qkv"), (768, 768, PK_EPI_BIAS_RESID_F32, "proj"),
                              (3072, 768, PK_EPI_BIAS_GELU_BF16, "fc1"), (768, 3072, PK_EPI_BIAS_RESID_F32, "fc2")]:
        a = (torch.randn(Mbig, K, device=DEV) * 0.5).to(torch.bfloat16)
        w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn(N, device=DEV) * 0.1
        bf = epi in (PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16)
        out = torch.empty(Mbig, N, device=DEV, dtype=torch.bfloat16 if bf else torch.float32)
        resid = None if bf else out
        for bn in (128, 256):
            ms = time_ms(lambda: ops.gemm(a, w, bias, out, epi, resid=resid, block_n=bn, cta_pair=2))
            report(f"pair_time_{name}_bn{bn}", ms=ms, tflops=2.0 * Mbig * N * K / ms / 1e9, flag=ops.device_flag())
        ms = time_ms(lambda: ops.gemm(a, w, bias, out, epi, resid=resid, block_n=256, cta_pair=1))
        report(f"single_time_{name}_bn256", ms=ms, tflops=2.0 * Mbig * N * K / ms / 1e9, flag=ops.device_flag())
        ms = time_ms(lambda: torch.nn.functional.linear(a, w))
        report(f"cublas_time_{name}", ms=ms, tflops=2.0 * Mbig * N * K / ms / 1e9)


def probe_attn2():
    """tcgen05/TMEM attention (uniform 128 < n <= 256, dh 64) against the fp32 expression and the general kernel."""
    torch.manual_seed(2)
    for (B, H, N) in [(1, 1, 197), (2, 12, 197), (3, 6, 198), (2, 3, 129), (2, 2, 256), (2, 2, 144), (5, 4, 145), (300, 12, 197)]:
        dh = 64
        D = H * dh
        qkv = (torch.randn(B * N, 3 * D, device=DEV) * (1.5 if B < 100 else 1.0)).to(torch.bfloat16)
        out = torch.full((B * N, D), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.attention(qkv, out, B, H, dh, seq_len=N, impl=2)
        flag = ops.device_flag()
        out1 = torch.zeros(B * N, D, device=DEV, dtype=torch.bfloat16)
        ops.attention(qkv, out1, B, H, dh, seq_len=N, impl=1)
        if B < 100:
            ref = ref_attention(qkv, B, H, dh, [N] * B)
        else:
            ref = out1.float()
        bad = ~((out.float() - ref).abs() <= 2e-2 * ref.abs().max())
        extra = {}
        if bad.any():
            rows = bad.any(1).nonzero().flatten()
            cols = bad.any(0).nonzero().flatten()
            extra = dict(bad_frac=bad.float().mean().item(), bad_rows=rows[:10].tolist(), n_bad_rows=int(rows.numel()),
                         bad_cols=cols[:10].tolist(), n_bad_cols=int(cols.numel()), nan=int(torch.isnan(out.float()).sum()),
                         got=out[rows[0], cols[0]].item(), ref=ref[rows[0], cols[0]].item())
        report(f"attn_tc_B{B}_H{H}_N{N}", err=rel_err(out, ref), err_general=rel_err(out1, ref), flag=flag, **extra)
        if flag:
            return
    B, H, N, dh = 256, 12, 197, 64
    D = H * dh
    qkv = torch.randn(B * N, 3 * D, device=DEV).to(torch.bfloat16)
    out = torch.zeros(B * N, D, device=DEV, dtype=torch.bfloat16)
    for impl in (1, 2):
        ms = time_ms(lambda: ops.attention(qkv, out, B, H, dh, seq_len=N, impl=impl))
        report(f"attn_time_impl{impl}_B{B}", ms=ms, tflops=4.0 * B * H * N * N * dh / ms / 1e9, flag=ops.device_flag())


def probe_ln():
    torch.manual_seed(1)
    for D in (64, 128, 192, 256, 384, 768, 1024):
        rows = 1237
        x = torch.randn(rows, D, device=DEV) * 2 + 0.3
        g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
        ref = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5)
        y = ops.layernorm(x, g, b, 1e-5)
        report(f"ln_D{D}", err=rel_err(y, ref))
    x = torch.randn(1000, 384, device=DEV)
    g, b = torch.randn(384, device=DEV), torch.randn(384, device=DEV)
    rs = torch.rand(500, device=DEV)
    idx = torch.randperm(1000, device=DEV)[:500].to(torch.int32)
    y = ops.layernorm(x, g, b, 1e-6, rowscale=rs, row_index=idx)
    ref = rs[:, None] * torch.nn.functional.layer_norm(x[idx.long()], (384,), g, b, 1e-6)
    report("ln_rowscale_index", err=rel_err(y, ref))
    rows = 197 * 2048
    x = torch.randn(rows, 768, device=DEV)
    g, b = torch.randn(768, device=DEV), torch.randn(768, device=DEV)
    y = torch.empty(rows, 768, device=DEV, dtype=torch.bfloat16)
    ms = time_ms(lambda: ops.layernorm(x, g, b, 1e-5, y))
    report("ln_time_vitb_2048", ms=ms, gbs=rows * 768 * 6 / ms / 1e6)


def ref_attention(qkv, B, H, dh, lens, key_mult=None, extra_kv=None, extra_mult=None):
    D = H * dh
    out = torch.zeros(qkv.shape[0], D, device=qkv.device)
    start = 0
    for b in range(B):
        n = lens[b]
        blk = qkv[start:start + n].float()
        q, k, v = blk[:, :D], blk[:, D:2 * D], blk[:, 2 * D:]
        q = q.view(n, H, dh).transpose(0, 1)
        k = k.view(n, H, dh).transpose(0, 1)
        v = v.view(n, H, dh).transpose(0, 1)
        bias = torch.zeros(n, device=qkv.device)
        if key_mult is not None:
            bias = key_mult[start:start + n].log()
        if extra_kv is not None and extra_mult[b] > 0:
            ek = extra_kv[:D].float().view(H, 1, dh)
            ev = extra_kv[D:].float().view(H, 1, dh)
            k = torch.cat([k, ek], 1)
            v = torch.cat([v, ev], 1)
            bias = torch.cat([bias, extra_mult[b:b + 1].log()])
        s = (q @ k.transpose(1, 2)) / math.sqrt(dh) + bias
        o = torch.softmax(s, -1) @ v
        out[start:start + n] = o.transpose(0, 1).reshape(n, D)
        start += n
    return out


def probe_attn():
    torch.manual_seed(2)
    for (B, H, dh, N) in [(3, 2, 64, 17), (2, 12, 64, 197), (2, 8, 32, 785), (4, 6, 64, 64), (1, 3, 64, 1)]:
        D = H * dh
        qkv = torch.randn(B * N, 3 * D, device=DEV).to(torch.bfloat16)
        out = torch.zeros(B * N, D, device=DEV, dtype=torch.bfloat16)
        ops.attention(qkv, out, B, H, dh, seq_len=N)
        report(f"attn_dense_B{B}_H{H}_dh{dh}_N{N}", err=rel_err(out, ref_attention(qkv, B, H, dh, [N] * B)))
    B, H, dh = 5, 6, 64
    D = H * dh
    lens = [3, 70, 198, 1, 129]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), device=DEV, dtype=torch.int32)
    rows = sum(lens)
    qkv = torch.randn(rows, 3 * D, device=DEV).to(torch.bfloat16)
    out = torch.zeros(rows, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, H, dh, cu_seqlens=cu, max_seq_len=max(lens))
    report("attn_ragged", err=rel_err(out, ref_attention(qkv, B, H, dh, lens)))
    km = torch.randint(1, 40, (rows,), device=DEV).float()
    ekv = (torch.randn(2 * D, device=DEV) * 0.5).to(torch.bfloat16)
    em = torch.tensor([0.0, 5.0, 100.0, 7.0, 0.0], device=DEV)
    ops.attention(qkv, out, B, H, dh, cu_seqlens=cu, max_seq_len=max(lens), key_mult=km, extra_kv=ekv, extra_mult=em)
    report("attn_ragged_mult_extra", err=rel_err(out, ref_attention(qkv, B, H, dh, lens, km, ekv, em)))
    B, H, dh, N = 256, 12, 64, 197
    D = H * dh
    qkv = torch.randn(B * N, 3 * D, device=DEV).to(torch.bfloat16)
    out = torch.zeros(B * N, D, device=DEV, dtype=torch.bfloat16)
    ms = time_ms(lambda: ops.attention(qkv, out, B, H, dh, seq_len=N))
    report("attn_time_vitb_256", ms=ms, tflops=4.0 * B * H * N * N * dh / ms / 1e9)


def probe_rows():
    torch.manual_seed(3)
    B, S, p, D = 3, 64, 8, 256
    img = torch.randn(B, 3, S, S, device=DEV)
    pt = ops.patchify(img, p)
    ref = torch.nn.functional.unfold(img, kernel_size=p, stride=p).transpose(1, 2).reshape(B * (S // p) ** 2, 3 * p * p)
    report("patchify_p8", err=rel_err(pt, ref.to(torch.bfloat16)))
    img = torch.randn(2, 3, 224, 224, device=DEV)
    pt = ops.patchify(img, 16)
    ref = torch.nn.functional.unfold(img, kernel_size=16, stride=16).transpose(1, 2).reshape(2 * 196, 768)
    report("patchify_p16", err=rel_err(pt, ref.to(torch.bfloat16)))
    seq, T = 20, 2
    x = torch.zeros(B * seq, D, device=DEV)
    tok, pos = torch.randn(T, D, device=DEV), torch.randn(seq, D, device=DEV)
    ops.fill_token_rows(x, B, seq, 1, tok, pos, scale=0.4)
    ref = torch.zeros(B, seq, D, device=DEV)
    ref[:, 1:1 + T] = 0.4 * tok + pos[1:1 + T]
    report("fill_token_rows", err=rel_err(x, ref.view(B * seq, D)))
    x.zero_()
    ops.fill_token_rows(x, B, seq, seq - 1, None, None, scale=0.7, n_tokens=1)
    report("fill_const_row", ok=bool((x.view(B, seq, D)[:, -1] == 0.7).all() and (x.view(B, seq, D)[:, :-1] == 0).all()))
    for (Bh, Dh, C, T) in [(5, 128, 10, 1), (19, 768, 1000, 1), (4, 384, 7, 2)]:
        seq = 11
        x = torch.randn(Bh * seq, Dh, device=DEV)
        g, b = torch.randn(Dh, device=DEV), torch.randn(Dh, device=DEV)
        hw, hb = torch.randn(C, Dh, device=DEV) / math.sqrt(Dh), torch.randn(C, device=DEV)
        got = ops.cls_head(x, Bh, seq, T, g, b, 1e-5, hw, hb)
        f = torch.nn.functional.layer_norm(x.view(Bh, seq, Dh)[:, :T], (Dh,), g, b, 1e-5).sum(1)
        report(f"cls_head_B{Bh}_D{Dh}_C{C}_T{T}", err=rel_err(got, f @ hw.t() + hb))


def probe_rank():
    torch.manual_seed(4)
    B, seq, D = 7, 197, 768
    x = torch.randn(B * seq, D, device=DEV)
    sc = ops.token_norm_score(x, B, seq)
    ref = torch.norm(x.view(B, seq, D)[:, 1:], dim=-1)
    report("token_norm_score", err=rel_err(sc, ref))
    for k in (1, 98, 196):
        kept = ops.topk_select(sc, k)
        exp = torch.argsort(sc, dim=-1, descending=True, stable=True)[:, :k]
        report(f"topk_k{k}", exact=bool(torch.equal(kept.long(), exp)))
    adv = torch.tensor([[1., 3, 3, 0, 3, 1, -0., 0.], [2.] * 8, [0., -0., 0., 1e-45, -0., 5, 5, 5]], device=DEV)
    for k in (1, 3, 8):
        kept = ops.topk_select(adv, k)
        exp = torch.argsort(adv.cpu(), dim=-1, descending=True, stable=True)[:, :k]
        report(f"topk_adversarial_k{k}", exact=bool(torch.equal(kept.cpu().long(), exp)), got=kept.cpu().tolist())
    big = torch.randn(3, 4096, device=DEV).round(decimals=1)      # many ties
    kept = ops.topk_select(big, 1000)
    exp = torch.argsort(big, dim=-1, descending=True, stable=True)[:, :1000]
    report("topk_ties_n4096", exact=bool(torch.equal(kept.long(), exp)))
    kept = ops.topk_select(sc, 98)
    y = ops.gather_rows(x, kept, B, seq)
    xr = x.view(B, seq, D)
    ref = torch.cat([xr[:, :1], torch.gather(xr[:, 1:], 1, kept.long().unsqueeze(-1).expand(-1, -1, D))], 1)
    report("gather_rows", exact=bool(torch.equal(y.view(B, 99, D), ref)))


def probe_model():
    import numpy as np
    from golden_cases import CASES, build_case
    from peekvit_b200.models import build_model
    from oracle import peekvit_oracle as po
    names = {"vit": "vit", "rankvit": "RankVisionTransformer"}
    for name, case in CASES.items():
        if case["family"] not in names:
            continue
        sd, images = build_case(case)
        model = build_model(names[case["family"]], case["cfg"])
        model.load_state_dict(sd, strict=True)
        model = model.to(DEV).eval()
        if case.get("budget") is not None:
            model.set_budget(case["budget"])
        aux = {}
        from peekvit_b200 import runner
        logits = runner.run(model, images.to(DEV), aux).cpu()
        ref = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))["logits"])
        report(f"model_{name}", err=rel_err(logits, ref), top1=float((logits.argmax(1) == ref.argmax(1)).float().mean()),
               flag=ops.device_flag(), seq_lens=aux.get("seq_lens"))
    # ViT-B/16 vs oracle on 8 images + throughput
    from oracle import weights as ow
    cfg = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
    sd = ow.make_state_dict("vit", cfg, seed=4321)
    images = ow.synthetic_images(8, 224, seed=1234)
    t0 = time.time()
    ref, _ = po.forward("vit", sd, cfg, images)
    cpu_s = time.time() - t0
    model = build_model("vit", cfg)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    logits = model(images.to(DEV)).cpu()
    report("model_vit_b16_8img", err=rel_err(logits, ref), top1=float((logits.argmax(1) == ref.argmax(1)).float().mean()),
           flag=ops.device_flag(), cpu_img_s=8 / cpu_s)
    big = torch.randn(1024, 3, 224, 224, device=DEV)
    for mb in (32, 64, 128, 256):
        model.pk_micro_batch = mb
        ms = time_ms(lambda: model(big), iters=3, warm=1)
        report(f"vit_b16_throughput_mb{mb}", ms=ms, img_s=1024 / ms * 1e3, flag=ops.device_flag())


def probe_sparse():
    """ResidualViT / AViT / MoE drop-ins against the reference fixtures + the oracle, with the
    per-layer diagnostics needed to debug compaction remotely."""
    import numpy as np
    from golden_cases import CASES, build_case
    from peekvit_b200.models import build_model
    from peekvit_b200 import runner
    from oracle import peekvit_oracle as po
    names = {"residualvit": "residualvit", "adavit": "adavit", "moevit": "vitmoe"}
    for name, case in CASES.items():
        fam = case["family"]
        if fam not in names:
            continue
        try:
            sd, images = build_case(case)
            model = build_model(names[fam], case["cfg"])
            model.load_state_dict(sd, strict=True)
            model = model.to(DEV).eval()
            if case.get("budget") is not None:
                model.set_budget(case["budget"])
            aux = {}
            logits = runner.run(model, images.to(DEV), aux).cpu()
            gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
            ref = torch.from_numpy(gold["logits"])
            info = dict(err=rel_err(logits, ref), flag=ops.device_flag(), nan=int(torch.isnan(logits).sum()))
            if fam == "residualvit":
                for i, blk in enumerate(model.encoder.layers):
                    if blk.mask is None:
                        continue
                    m, g = blk.mask.cpu(), torch.from_numpy(gold[f"mask_{i}"])
                    info[f"L{i}"] = dict(shape=list(m.shape), keep=float((m > 0).float().mean()), keep_ref=float((g > 0).float().mean()),
                                         agree=float(((m > 0) == (g > 0)).float().mean()), maxdiff=float((m - g).abs().max()),
                                         rows=[int(r) for r in aux["rows"][i]] if "rows" in aux else None)
            if fam == "adavit":
                info["rows"] = [int(r[0]) for r in aux.get("rows", [[]])[0]] if aux.get("rows") else None
                model.pk_early_exit = False
                logits2 = runner.run(model, images.to(DEV)).cpu()
                info["err_no_early_exit"] = rel_err(logits2, ref)
                cnt, gcnt = model.encoder.counter_token.cpu(), torch.from_numpy(gold["counter_token"])
                rho, grho = model.encoder.rho_token.cpu(), torch.from_numpy(gold["rho_token"])
                info["counter_agree"] = float((cnt == gcnt).float().mean())
                info["rho_maxdiff"] = float((rho - grho).abs().max())
            if fam == "moevit":
                for i, blk in enumerate(model.encoder.layers):
                    gp = blk.mlp.gating_probs
                    if gp is not None:
                        info[f"route_agree_L{i}"] = float((gp.argmax(-1).cpu().numpy() == gold[f"mlp_gating_{i}"]).mean())
            report(f"sparse_{name}", **info)
        except Exception as e:  # keep going: one broken family must not hide the others
            import traceback
            report(f"sparse_{name}", exception=repr(e), tb=traceback.format_exc()[-1500:])
            try:
                ops.device_flag()
            except Exception:
                pass


if __name__ == "__main__":
    which = sys.argv[1]
    out = None
    if "--json" in sys.argv:
        out = sys.argv[sys.argv.index("--json") + 1]
    print("device:", torch.cuda.get_device_name(0), flush=True)
    try:
        {"gemm": probe_gemm, "gemm2": probe_gemm2, "attn2": probe_attn2, "ln": probe_ln, "attn": probe_attn, "rows": probe_rows, "rank": probe_rank, "model": probe_model,
         "sparse": probe_sparse}[which]()
    finally:
        if out:
            os.makedirs(os.path.dirname(out), exist_ok=True)
            with open(out, "w") as f:
                json.dump(RESULTS, f, indent=1)
