"""Tensor-level wrappers over the C ABI (``include/peekvit_b200.h``).

Each function checks dtype / device / layout, then makes exactly one C call on torch's
current CUDA stream with raw device pointers.  PyTorch is used for memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib
from ._lib import (AttentionArgs, AvitArgs, CompactArgs, GemmArgs, ResidualGateArgs, PK_EPI_BIAS_BF16, PK_EPI_BIAS_F32, PK_EPI_BIAS_GELU_BF16,
                   PK_EPI_BIAS_RESID_F32, PK_OUT_BF16, PK_OUT_BF16X2, PK_OUT_F16, check)


# Launch accounting for bench.py: every wrapper below launches exactly one kernel of ours.
launch_count = 0
# When a list is installed here, gemm() brackets its launch with CUDA events on the launching
# stream and appends (start, end, flops): the live per-kernel timing the roofline line needs.
gemm_timeline = None


def _lib_for(t: torch.Tensor):
    global launch_count
    launch_count += 1
    if not t.is_cuda:
        raise RuntimeError("peekvit_b200 kernels need CUDA tensors on a B200; there is no CPU path")
    cur = torch.cuda.current_device()
    idx = t.device.index if t.device.index is not None else cur
    if idx != cur:
        # the library works on the context and the stream of the CURRENT device (runner.run / run_host enter
        # torch.cuda.device(model device) themselves); launching on another device's stream would be an illegal access
        raise RuntimeError(f"tensor on cuda:{idx} but the current device is cuda:{cur}: wrap the call in torch.cuda.device({idx})")
    return _lib.init(idx)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor], dtype=None) -> Optional[int]:
    if t is None:
        return None
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError("expected a CUDA tensor")
    return t.data_ptr()


def _rowmajor(t: torch.Tensor, name: str) -> int:
    """Leading dimension (elements) of a 2-D row-major tensor (last stride 1)."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be 2-D with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")
    return t.stride(0)


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, epilogue: int, *,
         resid: Optional[torch.Tensor] = None, rowscale: Optional[torch.Tensor] = None,
         rows_per_group: int = 0, group_stride: int = 0, group_offset: int = 0, resid_is_pos: bool = False,
         pos_offset: Optional[int] = None, m_dev: Optional[torch.Tensor] = None, row_begin_dev: Optional[torch.Tensor] = None,
         out_row_index: Optional[torch.Tensor] = None, block_n: int = 0, max_ctas: int = 0,
         m: Optional[int] = None, epilogue_mode: int = 0, cta_pair: int = 0,
         xb_out: Optional[torch.Tensor] = None, row_stats: Optional[torch.Tensor] = None,
         ln_stats: Optional[torch.Tensor] = None, ln_c1: Optional[torch.Tensor] = None, ln_dim: int = 0, ln_eps: float = 0.0,
         a_wrap_k: int = 0, out_format: int = PK_OUT_BF16, group_offsets: Optional[torch.Tensor] = None,
         n_groups: int = 0) -> torch.Tensor:
    """out = epilogue(a @ w.T + bias): a bf16 [M,K], w bf16 [N,K] (nn.Linear layout).  ``a_wrap_k`` = k: a is the two-term
    split row [lo | hi] (2k wide) read as lo, hi, hi against w = [Wh | Wl | Wh] (3k wide); ``out_format``: PK_OUT_F16 (half
    output) or PK_OUT_BF16X2 (out is [M, 2N]: the value split into [lo | hi])."""
    lib = _lib_for(a)
    lda, ldw, ldo = _rowmajor(a, "a"), _rowmajor(w, "w"), _rowmajor(out, "out")
    M = a.shape[0] if m is None else m
    K = 3 * a_wrap_k if a_wrap_k > 0 else a.shape[1]
    if a_wrap_k > 0 and a.shape[1] != 2 * a_wrap_k:
        raise ValueError(f"a_wrap_k={a_wrap_k} needs a [M, {2 * a_wrap_k}] split operand, got {tuple(a.shape)}")
    N = w.shape[0]
    if group_offsets is not None:                 # grouped launch: w is the groups' weights stacked, N one group's width
        if n_groups < 1 or N % n_groups != 0:
            raise ValueError(f"grouped GEMM: {N} stacked weight rows do not split into {n_groups} groups")
        N //= n_groups
    if w.shape[1] != K:
        raise ValueError(f"K mismatch: a {tuple(a.shape)} w {tuple(w.shape)}")
    out_dtype = torch.bfloat16 if epilogue in (PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16) else torch.float32
    if out_format == PK_OUT_F16:
        out_dtype = torch.float16
    if out_format == PK_OUT_BF16X2 and out.shape[1] != 2 * N:
        raise ValueError(f"split output needs out [M, {2 * N}], got {tuple(out.shape)}")
    args = GemmArgs()
    args.A, args.W = _ptr(a, torch.bfloat16), _ptr(w, torch.bfloat16)
    args.M, args.N, args.K = M, N, K
    args.lda, args.ldw = lda, ldw
    args.bias = _ptr(bias, torch.float32)
    args.epilogue = epilogue
    args.out, args.ldo = _ptr(out, out_dtype), ldo
    args.resid = _ptr(resid, torch.float32)
    args.ldr = _rowmajor(resid, "resid") if resid is not None else 0
    args.rowscale = _ptr(rowscale, torch.float32)
    args.rows_per_group, args.group_stride, args.group_offset = rows_per_group, group_stride, group_offset
    args.resid_is_pos = int(resid_is_pos)
    args.pos_offset = group_offset if pos_offset is None else pos_offset
    args.m_dev = _ptr(m_dev, torch.int32)
    args.row_begin_dev = _ptr(row_begin_dev, torch.int32)
    args.out_row_index = _ptr(out_row_index, torch.int32)
    args.block_n, args.max_ctas = block_n, max_ctas
    args.epilogue_mode = epilogue_mode
    args.cta_pair = cta_pair
    args.xb_out = _ptr(xb_out, torch.bfloat16)
    args.ldxb = _rowmajor(xb_out, "xb_out") if xb_out is not None else 0
    args.row_stats = _ptr(row_stats, torch.float32)
    args.ln_stats = _ptr(ln_stats, torch.float32)
    args.ln_c1 = _ptr(ln_c1, torch.float32)
    args.ln_parts = ln_stats.shape[1] if ln_stats is not None else 0
    args.ln_dim, args.ln_eps = ln_dim, ln_eps
    args.a_wrap_k, args.out_format = a_wrap_k, out_format
    args.group_offsets, args.n_groups = _ptr(group_offsets, torch.int32), n_groups
    if gemm_timeline is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.pk_gemm_bf16(C.byref(args), _stream()), "pk_gemm_bf16")
        e1.record()
        gemm_timeline.append((e0, e1, 2.0 * M * N * K))
        return out
    check(lib.pk_gemm_bf16(C.byref(args), _stream()), "pk_gemm_bf16")
    return out


def patchify(images: torch.Tensor, patch_size: int, out: Optional[torch.Tensor] = None, rows_per_sample: int = 0,
             row_offset: int = 0) -> torch.Tensor:
    lib = _lib_for(images)
    if images.dim() != 4 or images.shape[1] != 3 or images.shape[2] != images.shape[3] or not images.is_contiguous():
        raise ValueError(f"images must be contiguous [B,3,S,S], got {tuple(images.shape)}")
    B, _, S, _ = images.shape
    P, Kp = (S // patch_size) ** 2, 3 * patch_size * patch_size
    if out is None:
        out = torch.empty(B * P, Kp, dtype=torch.bfloat16, device=images.device)
    check(lib.pk_patchify(_ptr(images, torch.float32), _ptr(out, torch.bfloat16), B, S, patch_size, rows_per_sample, row_offset,
                          _stream()), "pk_patchify")
    return out


def fill_token_rows(x: torch.Tensor, batch: int, seq_stride: int, row_offset: int, tokens: Optional[torch.Tensor],
                    pos: Optional[torch.Tensor], scale: float = 1.0, n_tokens: Optional[int] = None,
                    pos_offset: Optional[int] = None) -> None:
    """``pos_offset``: first pos_embedding row to use when it differs from ``row_offset`` (ResidualViT keeps
    its budget token at local row 1, which shifts the other tokens but not their position rows)."""
    lib = _lib_for(x)
    dim = x.shape[-1]
    if tokens is not None:
        tokens = tokens.reshape(-1, dim)
        n_tokens = tokens.shape[0]
    pos_ptr = None
    if pos is not None:
        pos = pos.reshape(-1, dim)
        po = row_offset if pos_offset is None else pos_offset
        if po + n_tokens > pos.shape[0]:
            raise ValueError("pos_embedding has too few rows")
        # the kernel indexes pos[row_offset + t]; shift the base so that lands on pos[po + t]
        pos_ptr = _ptr(pos, torch.float32) + (po - row_offset) * dim * 4
    check(lib.pk_fill_token_rows(_ptr(x, torch.float32), batch, seq_stride, row_offset, n_tokens, dim,
                                 _ptr(tokens, torch.float32), pos_ptr, float(scale), _stream()),
          "pk_fill_token_rows")


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None, *,
              rows: Optional[int] = None, rowscale: Optional[torch.Tensor] = None, row_index: Optional[torch.Tensor] = None,
              rows_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib_for(x)
    dim = x.shape[-1]
    if rows is None:
        rows = row_index.numel() if row_index is not None else x.numel() // dim
    if out is None:
        out = torch.empty(rows, dim, dtype=torch.bfloat16, device=x.device)
    check(lib.pk_layernorm_bf16(_ptr(x, torch.float32), _ptr(out, torch.bfloat16), _ptr(gamma, torch.float32),
                                _ptr(beta, torch.float32), float(eps), rows, dim, _ptr(rowscale, torch.float32),
                                _ptr(row_index, torch.int32), _ptr(rows_dev, torch.int32), _stream()), "pk_layernorm_bf16")
    return out


def attention(qkv: torch.Tensor, out: torch.Tensor, batch: int, num_heads: int, head_dim: int, *, seq_len: int = 0,
              cu_seqlens: Optional[torch.Tensor] = None, max_seq_len: int = 0, key_mult: Optional[torch.Tensor] = None,
              extra_kv: Optional[torch.Tensor] = None, extra_mult: Optional[torch.Tensor] = None, impl: int = 0,
              half_split: bool = False, route_rows: Optional[torch.Tensor] = None, route_min_rows: int = 0) -> torch.Tensor:
    """``half_split`` (bf16x2 mode, tcgen05 kernel): qkv is IEEE half, out is bf16 [rows, 2*D] = the result split [lo | hi]."""
    lib = _lib_for(qkv)
    D = num_heads * head_dim
    if qkv.shape[-1] != 3 * D or not qkv.is_contiguous() or not out.is_contiguous() or out.shape[-1] != (2 * D if half_split else D):
        raise ValueError("qkv must be contiguous [rows, 3*D] and out contiguous [rows, D] ([rows, 2*D] for the split output)")
    a = AttentionArgs()
    a.qkv, a.out = _ptr(qkv, torch.float16 if half_split else torch.bfloat16), _ptr(out, torch.bfloat16)
    a.qkv_format, a.out_format = (PK_OUT_F16, PK_OUT_BF16X2) if half_split else (PK_OUT_BF16, PK_OUT_BF16)
    a.batch, a.num_heads, a.head_dim = batch, num_heads, head_dim
    a.seq_len = seq_len
    a.cu_seqlens = _ptr(cu_seqlens, torch.int32)
    a.max_seq_len = max_seq_len or seq_len
    a.scale = 1.0 / math.sqrt(head_dim)
    a.key_mult = _ptr(key_mult, torch.float32)
    a.extra_kv = _ptr(extra_kv, torch.bfloat16)
    a.extra_mult = _ptr(extra_mult, torch.float32)
    a.impl = impl
    a.total_rows = qkv.shape[0]
    a.route_rows, a.route_min_rows = _ptr(route_rows, torch.int32), int(route_min_rows)
    check(lib.pk_attention_fwd(C.byref(a), _stream()), "pk_attention_fwd")
    return out


def cls_head(x: torch.Tensor, batch: int, seq_len: int, n_cls: int, gamma, beta, eps: float, head_w, head_b,
             cu_seqlens: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib_for(x)
    dim = x.shape[-1]
    C_ = head_w.shape[0]
    if out is None:
        out = torch.empty(batch, C_, dtype=torch.float32, device=x.device)
    check(lib.pk_cls_head(_ptr(x, torch.float32), batch, seq_len, _ptr(cu_seqlens, torch.int32), n_cls, dim,
                          _ptr(gamma, torch.float32), _ptr(beta, torch.float32), float(eps),
                          _ptr(head_w, torch.float32), _ptr(head_b, torch.float32), C_, _ptr(out, torch.float32), _stream()),
          "pk_cls_head")
    return out


def cls_features(x: torch.Tensor, batch: int, seq_len: int, n_cls: int, gamma, beta, eps: float, out: torch.Tensor,
                 cu_seqlens: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b] = sum over the class tokens of LN(x[class rows of sample b]) (f32 [batch, dim])."""
    lib = _lib_for(x)
    check(lib.pk_cls_features(_ptr(x, torch.float32), batch, seq_len, _ptr(cu_seqlens, torch.int32), n_cls, x.shape[-1],
                              _ptr(gamma, torch.float32), _ptr(beta, torch.float32), float(eps), _ptr(out, torch.float32), _stream()),
          "pk_cls_features")
    return out


def argmax_count(logits: torch.Tensor, labels: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None,
                 pred: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Top-1 prediction (int32) and/or accuracy counts (int64 [correct, total], accumulated) on the device."""
    lib = _lib_for(logits)
    if logits.dim() != 2 or not logits.is_contiguous():
        raise ValueError("logits must be contiguous [batch, classes]")
    B, C_ = logits.shape
    if pred is None and counts is None:
        pred = torch.empty(B, dtype=torch.int32, device=logits.device)
    check(lib.pk_argmax_count(_ptr(logits, torch.float32), _ptr(labels, torch.int64), B, C_, _ptr(pred, torch.int32),
                              _ptr(counts, torch.int64), _stream()), "pk_argmax_count")
    return pred if pred is not None else counts


def token_norm_score(x: torch.Tensor, batch: int, seq_len: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib_for(x)
    dim = x.shape[-1]
    if out is None:
        out = torch.empty(batch, seq_len - 1, dtype=torch.float32, device=x.device)
    check(lib.pk_token_norm_score(_ptr(x, torch.float32), _ptr(out, torch.float32), batch, seq_len, dim, _stream()),
          "pk_token_norm_score")
    return out


def topk_select(scores: torch.Tensor, k: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib_for(scores)
    batch, n = scores.shape
    if not scores.is_contiguous():
        raise ValueError("scores must be contiguous")
    if out is None:
        out = torch.empty(batch, k, dtype=torch.int32, device=scores.device)
    check(lib.pk_topk_select(_ptr(scores, torch.float32), _ptr(out, torch.int32), batch, n, k, _stream()), "pk_topk_select")
    return out


def gather_rows(x: torch.Tensor, kept: torch.Tensor, batch: int, seq_len: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib_for(x)
    dim = x.shape[-1]
    k = kept.shape[1]
    if out is None:
        out = torch.empty(batch * (k + 1), dim, dtype=torch.float32, device=x.device)
    check(lib.pk_gather_rows(_ptr(x, torch.float32), _ptr(out, torch.float32), _ptr(kept, torch.int32), batch, seq_len, k, dim,
                             _stream()), "pk_gather_rows")
    return out


def exclusive_scan(lens: torch.Tensor, cu_out: torch.Tensor, total_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib_for(lens)
    n = lens.numel()
    check(lib.pk_exclusive_scan_i32(_ptr(lens, torch.int32), n, _ptr(cu_out, torch.int32), _ptr(total_out, torch.int32), _stream()),
          "pk_exclusive_scan_i32")
    return cu_out


def compact_rows(x_in, x_out, cu_in, cu_out, batch: int, rows_in_cap: int, dst_local, sample_of, *, scale_in=None, scale_out=None,
                 attrs=(), ghost: bool = False, publish=None) -> None:
    """attrs: up to three (in, out) pairs of per-row fp32 attributes.  publish = (tok_row, mask_pub, n_img): the same launch
    also does ``residual_publish`` with ``scale_in`` as the soft mask."""
    lib = _lib_for(x_in)
    a = CompactArgs()
    a.x_in, a.x_out, a.dim = _ptr(x_in, torch.float32), _ptr(x_out, torch.float32), x_in.shape[-1]
    a.cu_in, a.cu_out, a.batch = _ptr(cu_in, torch.int32), _ptr(cu_out, torch.int32), batch
    a.rows_in_cap = rows_in_cap
    a.dst_local, a.sample_of = _ptr(dst_local, torch.int32), _ptr(sample_of, torch.int32)
    a.scale_in, a.scale_out = _ptr(scale_in, torch.float32), _ptr(scale_out, torch.float32)
    pairs = list(attrs) + [(None, None)] * (3 - len(attrs))
    a.a0_in, a.a0_out = _ptr(pairs[0][0], torch.float32), _ptr(pairs[0][1], torch.float32)
    a.a1_in, a.a1_out = _ptr(pairs[1][0], torch.float32), _ptr(pairs[1][1], torch.float32)
    a.a2_in, a.a2_out = _ptr(pairs[2][0], torch.float32), _ptr(pairs[2][1], torch.float32)
    a.ghost = int(ghost)
    if publish is not None:
        if scale_in is None:
            raise ValueError("compact_rows: the fused publish reads the soft mask from scale_in")
        a.pub_tok_row, a.pub_mask, a.pub_n_img = _ptr(publish[0], torch.int32), _ptr(publish[1], torch.float32), int(publish[2])
    check(lib.pk_compact_rows(C.byref(a), _stream()), "pk_compact_rows")


def budget_mean_threshold(x, cu, batch: int, budget_pos: int, thr_out) -> None:
    lib = _lib_for(x)
    check(lib.pk_budget_mean_threshold(_ptr(x, torch.float32), _ptr(cu, torch.int32), batch, budget_pos, x.shape[-1],
                                       _ptr(thr_out, torch.float32), _stream()), "pk_budget_mean_threshold")


def residual_gate_plan(x, cu_in, mult_in, batch: int, max_seq_len: int, *, n_special: int, budget_pos: int, gated: bool,
                       gate_w=None, gate_b: float = 0.0, gate_temp: float = 1.0, gate_bias: float = 0.0, gate_type: int = 0,
                       thr_mode: int = 2, bt_w=None, bt_b: float = 0.0, thr_dev=None, thr_const: float = 0.5,
                       mask, dst_local, sample_of, new_len, mdrop) -> None:
    lib = _lib_for(x)
    a = ResidualGateArgs()
    a.x, a.cu_in, a.mult_in = _ptr(x, torch.float32), _ptr(cu_in, torch.int32), _ptr(mult_in, torch.float32)
    a.batch, a.dim, a.max_seq_len = batch, x.shape[-1], max_seq_len
    a.n_special, a.budget_pos, a.gated = n_special, budget_pos, int(gated)
    a.gate_w, a.gate_b, a.gate_temp, a.gate_bias, a.gate_type = _ptr(gate_w, torch.float32), gate_b, gate_temp, gate_bias, gate_type
    a.thr_mode, a.bt_w, a.bt_b = thr_mode, _ptr(bt_w, torch.float32), bt_b
    a.thr_dev, a.thr_const = _ptr(thr_dev, torch.float32), thr_const
    a.mask, a.dst_local, a.sample_of = _ptr(mask, torch.float32), _ptr(dst_local, torch.int32), _ptr(sample_of, torch.int32)
    a.new_len, a.mdrop = _ptr(new_len, torch.int32), _ptr(mdrop, torch.float32)
    check(lib.pk_residual_gate_plan(C.byref(a), _stream()), "pk_residual_gate_plan")


def residual_ghost(x, mult, cu, mdrop, mlp0, batch: int) -> None:
    lib = _lib_for(x)
    check(lib.pk_residual_ghost(_ptr(x, torch.float32), _ptr(mult, torch.float32), _ptr(cu, torch.int32), _ptr(mdrop, torch.float32),
                                _ptr(mlp0, torch.float32), batch, x.shape[-1], _stream()), "pk_residual_ghost")


def residual_publish(mask, dst_local, cu_out, tok_row, mask_pub, batch: int, n_img: int) -> None:
    lib = _lib_for(mask)
    check(lib.pk_residual_publish(_ptr(mask, torch.float32), _ptr(dst_local, torch.int32), _ptr(cu_out, torch.int32),
                                  _ptr(tok_row, torch.int32), _ptr(mask_pub, torch.float32), batch, n_img, _stream()),
          "pk_residual_publish")


def avit_halt_plan(x, cu_in, batch: int, seq_total: int, c, R, tokid, *, gate_scale: float, gate_center: float, eps: float,
                   last_layer: bool, early_exit: bool, out_acc, rho=None, counter=None, dst_local, sample_of, new_len, n_halted) -> None:
    lib = _lib_for(x)
    a = AvitArgs()
    a.x, a.cu_in, a.batch, a.dim, a.seq_total = _ptr(x, torch.float32), _ptr(cu_in, torch.int32), batch, x.shape[-1], seq_total
    a.c, a.R, a.tokid = _ptr(c, torch.float32), _ptr(R, torch.float32), _ptr(tokid, torch.float32)
    a.gate_scale, a.gate_center, a.eps = gate_scale, gate_center, eps
    a.last_layer, a.early_exit = int(last_layer), int(early_exit)
    a.out_acc, a.rho, a.counter = _ptr(out_acc, torch.float32), _ptr(rho, torch.float32), _ptr(counter, torch.float32)
    a.dst_local, a.sample_of = _ptr(dst_local, torch.int32), _ptr(sample_of, torch.int32)
    a.new_len, a.n_halted = _ptr(new_len, torch.int32), _ptr(n_halted, torch.float32)
    check(lib.pk_avit_halt_plan(C.byref(a), _stream()), "pk_avit_halt_plan")


MOE_SORT_SCRATCH_INTS = 4096      # PK_MOE_SORT_SCRATCH_INTS


def moe_route(x, gamma, beta, eps: float, gate_w, gate_b, rows: int, expert, offsets, counts, src_of, scratch=None) -> None:
    lib = _lib_for(x)
    if scratch is None:
        scratch = torch.empty(MOE_SORT_SCRATCH_INTS, dtype=torch.int32, device=x.device)
    check(lib.pk_moe_route(_ptr(x, torch.float32), _ptr(gamma, torch.float32), _ptr(beta, torch.float32), float(eps),
                           _ptr(gate_w, torch.float32), _ptr(gate_b, torch.float32), gate_w.shape[0], rows, x.shape[-1],
                           _ptr(expert, torch.int32), _ptr(offsets, torch.int32), _ptr(counts, torch.int32),
                           _ptr(src_of, torch.int32), _ptr(scratch, torch.int32), _stream()), "pk_moe_route")


def device_flag(reset: bool = True) -> int:
    """Watchdog word of the tcgen05 kernels (0 = healthy) of the current device. Synchronises the device."""
    return _lib.init(torch.cuda.current_device()).pk_device_flag(int(reset))


_flag_mirrors = {}


def device_flag_mirror(device_index: int) -> torch.Tensor:
    """Pinned host word that ``device_flag_async`` refreshes in stream order (one per device)."""
    m = _flag_mirrors.get(device_index)
    if m is None:
        m = _flag_mirrors[device_index] = torch.zeros(1, dtype=torch.int32, pin_memory=True)
    return m


def device_flag_async(device_index: int) -> None:
    """Enqueue a copy of the watchdog word into its pinned host mirror on the current stream (no synchronisation)."""
    lib = _lib.init(device_index)
    check(lib.pk_device_flag_async(device_flag_mirror(device_index).data_ptr(), _stream()), "pk_device_flag_async")


def raise_if_flagged(device_index: int, sync: bool = False) -> None:
    """An expired bounded mbarrier wait makes the tcgen05 kernels bail out and leaves partial outputs: tell the caller.
    ``sync`` reads the device word itself (after a synchronisation the caller has done anyway); otherwise the host mirror,
    i.e. the state as of the last completed forward."""
    if sync:
        code = device_flag(reset=True)
        device_flag_mirror(device_index).zero_()
    else:
        m = device_flag_mirror(device_index)
        code = int(m[0])
        if code != 0:
            device_flag(reset=True)
            m.zero_()
    if code != 0:
        raise _lib.PkError(f"peekvit_b200 watchdog: a bounded wait expired inside a tcgen05 kernel (code {code & 0xffffffff:#x}); "
                           "the outputs produced since the previous check are invalid")


def scatter_add_rows(x: torch.Tensor, y: torch.Tensor, src_of: torch.Tensor, rows: Optional[int] = None) -> torch.Tensor:
    """x[src_of[r]] += y[r] (un-permute + residual add of expert-sorted rows)."""
    lib = _lib_for(x)
    n = y.shape[0] if rows is None else rows
    check(lib.pk_scatter_add_rows(_ptr(x, torch.float32), _ptr(y, torch.float32), _ptr(src_of, torch.int32), n, x.shape[-1], _stream()),
          "pk_scatter_add_rows")
    return x


def expert_onehot(expert: torch.Tensor, n_experts: int, out: torch.Tensor, rows: Optional[int] = None) -> torch.Tensor:
    """out[e, r] = (expert[r] == e) as f32."""
    lib = _lib_for(expert)
    n = expert.shape[0] if rows is None else rows
    check(lib.pk_expert_onehot(_ptr(expert, torch.int32), _ptr(out, torch.float32), n, n_experts, _stream()), "pk_expert_onehot")
    return out


def noise_snr(x: torch.Tensor, noise: torch.Tensor, snr_db: float, rows: Optional[int] = None) -> torch.Tensor:
    """x[r] += noise[r] * sqrt(mean(x[r]^2) / 10^(snr_db/10)), in place."""
    lib = _lib_for(x)
    n = x.shape[0] if rows is None else rows
    check(lib.pk_noise_snr(_ptr(x, torch.float32), _ptr(noise, torch.float32), n, x.shape[-1], float(snr_db), _stream()), "pk_noise_snr")
    return x


def zero_token_rows(x: torch.Tensor, batch: int, seq: int, tokens: torch.Tensor) -> torch.Tensor:
    """x[b*seq + tokens[j]] = 0 for every sample b, in place."""
    lib = _lib_for(x)
    check(lib.pk_zero_token_rows(_ptr(x, torch.float32), batch, seq, _ptr(tokens, torch.int32), tokens.numel(), x.shape[-1], _stream()),
          "pk_zero_token_rows")
    return x


def row_scale_add(out: torch.Tensor, a: torch.Tensor, scale: torch.Tensor, rows: int, *, one_minus: bool = False,
                  accumulate: bool = False) -> torch.Tensor:
    """out[r] = (out[r] if accumulate else 0) + w[r] * a[r], w = scale or 1 - scale (residualvit.py:145,176,239-242)."""
    lib = _lib_for(a)
    check(lib.pk_row_scale_add(_ptr(out, torch.float32), _ptr(a, torch.float32), _ptr(scale, torch.float32), rows, a.shape[-1],
                               int(one_minus), int(accumulate), _stream()), "pk_row_scale_add")
    return out


SPLIT_NONE, SPLIT_GELU, SPLIT_LAYERNORM = 0, 1, 2


def split3(x: torch.Tensor, out: torch.Tensor, mode: int = SPLIT_NONE, gamma: Optional[torch.Tensor] = None,
           beta: Optional[torch.Tensor] = None, eps: float = 0.0, rows: Optional[int] = None, rowscale: Optional[torch.Tensor] = None,
           row_index: Optional[torch.Tensor] = None, rows_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 rows -> bf16 split activation rows [m|l|h|m|h|h] (out [rows, 6*dim]), optionally after exact GELU / LayerNorm
    (times ``rowscale``; LayerNorm mode may gather its input rows through ``row_index``)."""
    lib = _lib_for(x)
    n = x.shape[0] if rows is None else rows
    check(lib.pk_split3_bf16(_ptr(x, torch.float32), _ptr(out, torch.bfloat16), n, x.shape[-1], mode, _ptr(gamma, torch.float32),
                             _ptr(beta, torch.float32), float(eps), _ptr(rowscale, torch.float32), _ptr(row_index, torch.int32),
                             _ptr(rows_dev, torch.int32), _stream()), "pk_split3_bf16")
    return out


def split(x: torch.Tensor, out: torch.Tensor, terms: int, mode: int = SPLIT_NONE, gamma: Optional[torch.Tensor] = None,
          beta: Optional[torch.Tensor] = None, eps: float = 0.0, rows: Optional[int] = None, rowscale: Optional[torch.Tensor] = None,
          row_index: Optional[torch.Tensor] = None, rows_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``split3`` (terms == 3: out [rows, 6*dim]) or its two-term form (terms == 2: out [rows, 2*dim] = [lo | hi], read by
    ``gemm(..., a_wrap_k=dim)``)."""
    if terms == 3:
        return split3(x, out, mode, gamma, beta, eps, rows, rowscale, row_index, rows_dev)
    lib = _lib_for(x)
    n = x.shape[0] if rows is None else rows
    check(lib.pk_split2_bf16(_ptr(x, torch.float32), _ptr(out, torch.bfloat16), n, x.shape[-1], mode, _ptr(gamma, torch.float32),
                             _ptr(beta, torch.float32), float(eps), _ptr(rowscale, torch.float32), _ptr(row_index, torch.int32),
                             _ptr(rows_dev, torch.int32), _stream()), "pk_split2_bf16")
    return out


def split_weight(w: torch.Tensor, terms: int) -> torch.Tensor:
    """Host-side prepack of an fp32 ``[N, K]`` weight for the split-operand GEMMs: three terms -> ``split3_weight``; two terms
    -> [Wh | Wl | Wh] (bf16 [N, 3K]) against activations read as lo, hi, hi: the products lo*Wh, hi*Wl, hi*Wh, small first."""
    if terms == 3:
        return split3_weight(w)
    w = w.detach().float()
    h = w.to(torch.bfloat16)
    l = (w - h.float()).to(torch.bfloat16)
    return torch.cat([h, l, h], dim=1).contiguous()


def split_width(terms: int) -> int:
    """Width of a split activation row in units of the unsplit width (the GEMM's A operand)."""
    return 6 if terms == 3 else 2


def patchify_split(images: torch.Tensor, patch_size: int, out: torch.Tensor, terms: int) -> torch.Tensor:
    """im2col straight into split rows: [m|l|h|m|h|h] (terms 3, 6*Kp wide) or [lo|hi|hi] (terms 2, 3*Kp wide, no wrap)."""
    if terms == 3:
        return patchify_split3(images, patch_size, out)
    lib = _lib_for(images)
    B, _, S, _ = images.shape
    check(lib.pk_patchify_split2(_ptr(images, torch.float32), _ptr(out, torch.bfloat16), B, S, patch_size, _stream()), "pk_patchify_split2")
    return out


def split3_weight(w: torch.Tensor) -> torch.Tensor:
    """Host-side prepack of an fp32 ``[N, K]`` weight into the split row [m|h|l|h|m|h] (bf16 [N, 6K]) that pairs with
    ``split3`` activations: the six K-wide products are mm, lh, hl, mh, hm, hh."""
    w = w.detach().float()
    h = w.to(torch.bfloat16)
    r = w - h.float()
    m = r.to(torch.bfloat16)
    l = (r - m.float()).to(torch.bfloat16)
    return torch.cat([m, h, l, h, m, h], dim=1).contiguous()


def patchify_split3(images: torch.Tensor, patch_size: int, out: torch.Tensor) -> torch.Tensor:
    lib = _lib_for(images)
    B, _, S, _ = images.shape
    check(lib.pk_patchify_split3(_ptr(images, torch.float32), _ptr(out, torch.bfloat16), B, S, patch_size, _stream()), "pk_patchify_split3")
    return out


def attention_f32(qkv: torch.Tensor, out: torch.Tensor, batch: int, num_heads: int, head_dim: int, seq_len: int = 0, *,
                  cu_seqlens: Optional[torch.Tensor] = None, max_seq_len: int = 0, key_mult: Optional[torch.Tensor] = None,
                  extra_kv: Optional[torch.Tensor] = None, extra_mult: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 attention core; same sample / multiplicity / virtual-key arguments as ``attention`` (extra_kv fp32 [2*D])."""
    lib = _lib_for(qkv)
    check(lib.pk_attention_f32(_ptr(qkv, torch.float32), _ptr(out, torch.float32), batch, num_heads, head_dim, max_seq_len or seq_len,
                               float(head_dim) ** -0.5, _ptr(cu_seqlens, torch.int32), _ptr(key_mult, torch.float32),
                               _ptr(extra_kv, torch.float32), _ptr(extra_mult, torch.float32), _stream()), "pk_attention_f32")
    return out


def gemm_row_stat_parts(n: int) -> int:
    """Statistics slots per row written by the LayerNorm-producer epilogue for an n-column output."""
    return int(_lib.load().pk_gemm_row_stat_parts(n))


def row_stats_cast(x: torch.Tensor, xb: torch.Tensor, stats: torch.Tensor, rows: Optional[int] = None) -> None:
    """xb = bf16(x), stats[r][0] = (sum, sum of squares) of row r, other slots zero."""
    lib = _lib_for(x)
    n = x.shape[0] if rows is None else rows
    check(lib.pk_row_stats_cast(_ptr(x, torch.float32), _ptr(xb, torch.bfloat16), _ptr(stats, torch.float32), n, x.shape[-1],
                                stats.shape[1], _stream()), "pk_row_stats_cast")


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)      # reference data/imagenette.py:72


def patchify_u8(images_hwc: torch.Tensor, patch_size: int, out: Optional[torch.Tensor] = None, mean=IMAGENET_MEAN,
                std=IMAGENET_STD, rows_per_sample: int = 0, row_offset: int = 0) -> torch.Tensor:
    """uint8 [B,S,S,3] -> normalised bf16 patches [B*(S/p)^2, 3*p*p] (ToTensor + Normalize + im2col in one pass)."""
    lib = _lib_for(images_hwc)
    if images_hwc.dim() != 4 or images_hwc.shape[3] != 3 or images_hwc.shape[1] != images_hwc.shape[2] or not images_hwc.is_contiguous():
        raise ValueError("images must be contiguous uint8 [B, S, S, 3]")
    B, S = images_hwc.shape[0], images_hwc.shape[1]
    P, Kp = (S // patch_size) ** 2, 3 * patch_size * patch_size
    if out is None:
        out = torch.empty(B * P, Kp, dtype=torch.bfloat16, device=images_hwc.device)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    check(lib.pk_patchify_u8(_ptr(images_hwc, torch.uint8), _ptr(out, torch.bfloat16), B, S, patch_size, m, s, rows_per_sample, row_offset,
                             _stream()), "pk_patchify_u8")
    return out


# ---------------------------------------------------------------- fine-tuning path (csrc/pk_train.cu)
def cast_bf16(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    lib = _lib_for(x)
    check(lib.pk_cast_f32_bf16(_ptr(x, torch.float32), _ptr(out, torch.bfloat16), x.numel(), _stream()), "pk_cast_f32_bf16")
    return out


def gelu_bf16(h_pre: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    lib = _lib_for(h_pre)
    check(lib.pk_gelu_bf16(_ptr(h_pre, torch.bfloat16), _ptr(out, torch.bfloat16), h_pre.numel(), _stream()), "pk_gelu_bf16")
    return out


def gelu_bwd_bf16(h_pre: torch.Tensor, dhid: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    lib = _lib_for(h_pre)
    check(lib.pk_gelu_bwd_bf16(_ptr(h_pre, torch.bfloat16), _ptr(dhid, torch.bfloat16), _ptr(out, torch.bfloat16), h_pre.numel(), _stream()),
          "pk_gelu_bwd_bf16")
    return out


def layernorm_bwd(x: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, eps: float, dx: torch.Tensor, rows: int, *,
                  row_index: Optional[torch.Tensor] = None, dy_div: int = 1, accumulate: bool = True) -> torch.Tensor:
    lib = _lib_for(x)
    check(lib.pk_layernorm_bwd(_ptr(x, torch.float32), _ptr(dy, torch.float32), _ptr(gamma, torch.float32), float(eps),
                               _ptr(dx, torch.float32), rows, x.shape[-1], _ptr(row_index, torch.int32), dy_div, int(accumulate),
                               _stream()), "pk_layernorm_bwd")
    return dx


def layernorm_bwd_gated(x: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, dx: torch.Tensor,
                        rows: int, rowscale: torch.Tensor, dot_out: torch.Tensor, *, accumulate: bool = True) -> torch.Tensor:
    """dx (+)= LN_bwd(rowscale * dy); dot_out[r] += dy[r] . LN(x)[r]  (a LayerNorm whose output is masked, residualvit.py:252,258)."""
    lib = _lib_for(x)
    check(lib.pk_layernorm_bwd_gated(_ptr(x, torch.float32), _ptr(dy, torch.float32), _ptr(gamma, torch.float32), _ptr(beta, torch.float32),
                                     float(eps), _ptr(dx, torch.float32), rows, x.shape[-1], _ptr(rowscale, torch.float32),
                                     _ptr(dot_out, torch.float32), int(accumulate), _stream()), "pk_layernorm_bwd_gated")
    return dx


def cast_rows_bf16(x: torch.Tensor, out: torch.Tensor, rowscale: torch.Tensor, rows: int) -> torch.Tensor:
    lib = _lib_for(x)
    check(lib.pk_cast_rows_f32_bf16(_ptr(x, torch.float32), _ptr(out, torch.bfloat16), _ptr(rowscale, torch.float32), rows, x.shape[-1],
                                    _stream()), "pk_cast_rows_f32_bf16")
    return out


def rowdot(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, rows: int, *, c: Optional[torch.Tensor] = None,
           div: Optional[torch.Tensor] = None, alpha: float = 1.0, accumulate: bool = True) -> torch.Tensor:
    """out[r] (+)= alpha * sum_d a[r,d] * (b[r,d] - c[r,d]) / div[r]; rows with div[r] <= 0 contribute nothing."""
    lib = _lib_for(a)
    check(lib.pk_rowdot(_ptr(a, torch.float32), _ptr(b, torch.float32), _ptr(c, torch.float32), _ptr(div, torch.float32),
                        _ptr(out, torch.float32), rows, a.shape[-1], float(alpha), int(accumulate), _stream()), "pk_rowdot")
    return out


def residual_gate_train_fwd(x: torch.Tensor, batch: int, seq: int, n_special: int, budget_pos: int, gate_w: torch.Tensor,
                            gate_b: torch.Tensor, gate_temp: float, gate_bias: float, bt_w: torch.Tensor, bt_b: torch.Tensor, rowscale: torch.Tensor,
                            mask: torch.Tensor, sig: torch.Tensor, thr: torch.Tensor) -> None:
    lib = _lib_for(x)
    check(lib.pk_residual_gate_train_fwd(_ptr(x, torch.float32), batch, seq, n_special, budget_pos, x.shape[-1], _ptr(gate_w, torch.float32),
                                         _ptr(gate_b, torch.float32), float(gate_temp), float(gate_bias), _ptr(bt_w, torch.float32),
                                         _ptr(bt_b, torch.float32),
                                         _ptr(rowscale, torch.float32), _ptr(mask, torch.float32), _ptr(sig, torch.float32),
                                         _ptr(thr, torch.float32), _stream()), "pk_residual_gate_train_fwd")


def residual_gate_train_bwd(x: torch.Tensor, dm: torch.Tensor, dmask_ext: Optional[torch.Tensor], mask: torch.Tensor, sig: torch.Tensor,
                            thr: torch.Tensor, batch: int, seq: int, n_special: int, budget_pos: int, gate_w: torch.Tensor,
                            gate_temp: float, bt_w: torch.Tensor, dx: torch.Tensor, g_gate_w: torch.Tensor, g_gate_b: torch.Tensor,
                            g_bt_w: torch.Tensor, g_bt_b: torch.Tensor) -> None:
    lib = _lib_for(x)
    check(lib.pk_residual_gate_train_bwd(_ptr(x, torch.float32), _ptr(dm, torch.float32), _ptr(dmask_ext, torch.float32),
                                         _ptr(mask, torch.float32), _ptr(sig, torch.float32), _ptr(thr, torch.float32), batch, seq,
                                         n_special, budget_pos, x.shape[-1], _ptr(gate_w, torch.float32), float(gate_temp),
                                         _ptr(bt_w, torch.float32), _ptr(dx, torch.float32), _ptr(g_gate_w, torch.float32),
                                         _ptr(g_gate_b, torch.float32), _ptr(g_bt_w, torch.float32), _ptr(g_bt_b, torch.float32),
                                         _stream()), "pk_residual_gate_train_bwd")


def attention_bwd(qkv: torch.Tensor, out: torch.Tensor, dout: torch.Tensor, dqkv: torch.Tensor, batch: int, num_heads: int,
                  head_dim: int, seq_len: int) -> torch.Tensor:
    lib = _lib_for(qkv)
    check(lib.pk_attention_bwd(_ptr(qkv, torch.bfloat16), _ptr(out, torch.bfloat16), _ptr(dout, torch.bfloat16),
                               _ptr(dqkv, torch.bfloat16), batch, num_heads, head_dim, seq_len, 1.0 / math.sqrt(head_dim), _stream()),
          "pk_attention_bwd")
    return dqkv


def softmax_xent(logits: torch.Tensor, labels: torch.Tensor, inv_count: float, loss_sum: torch.Tensor, dlogits: torch.Tensor,
                 correct: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib_for(logits)
    check(lib.pk_softmax_xent(_ptr(logits, torch.float32), _ptr(labels, torch.int64), logits.shape[0], logits.shape[1], float(inv_count),
                              _ptr(loss_sum, torch.float32), _ptr(dlogits, torch.float32), _ptr(correct, torch.int32), _stream()),
          "pk_softmax_xent")
    return dlogits


def head_bwd(dlogits: torch.Tensor, feat: torch.Tensor, weight: torch.Tensor, d_weight: torch.Tensor, d_bias: torch.Tensor,
             d_feat: torch.Tensor) -> None:
    lib = _lib_for(dlogits)
    B, C_ = dlogits.shape
    check(lib.pk_head_bwd(_ptr(dlogits, torch.float32), _ptr(feat, torch.float32), _ptr(weight, torch.float32), B, C_, feat.shape[1],
                          _ptr(d_weight, torch.float32), _ptr(d_bias, torch.float32), _ptr(d_feat, torch.float32), _stream()), "pk_head_bwd")


def sum_token_rows(x: torch.Tensor, batch: int, seq: int, row0: int, n_rows: int, out: torch.Tensor) -> torch.Tensor:
    lib = _lib_for(x)
    check(lib.pk_sum_token_rows(_ptr(x, torch.float32), batch, seq, row0, n_rows, x.shape[-1], _ptr(out, torch.float32), _stream()),
          "pk_sum_token_rows")
    return out


def scatter_rows(y: torch.Tensor, x: torch.Tensor, kept: torch.Tensor, batch: int, seq_len: int) -> torch.Tensor:
    """Backward of gather_rows: x[b*seq_len + tok] = y rows (x zeroed by the caller)."""
    lib = _lib_for(y)
    check(lib.pk_scatter_rows(_ptr(y, torch.float32), _ptr(x, torch.float32), _ptr(kept, torch.int32), batch, seq_len, kept.shape[1],
                              y.shape[-1], _stream()), "pk_scatter_rows")
    return x
