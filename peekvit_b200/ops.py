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
from ._lib import (AttentionArgs, GemmArgs, PK_EPI_BIAS_BF16, PK_EPI_BIAS_F32, PK_EPI_BIAS_GELU_BF16,
                   PK_EPI_BIAS_RESID_F32, check)


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
    return _lib.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


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
         m_dev: Optional[torch.Tensor] = None, block_n: int = 0, max_ctas: int = 0, m: Optional[int] = None) -> torch.Tensor:
    """out = epilogue(a @ w.T + bias): a bf16 [M,K], w bf16 [N,K] (nn.Linear layout)."""
    lib = _lib_for(a)
    lda, ldw, ldo = _rowmajor(a, "a"), _rowmajor(w, "w"), _rowmajor(out, "out")
    M, K = a.shape if m is None else (m, a.shape[1])
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"K mismatch: a {tuple(a.shape)} w {tuple(w.shape)}")
    out_dtype = torch.bfloat16 if epilogue in (PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16) else torch.float32
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
    args.m_dev = _ptr(m_dev, torch.int32)
    args.block_n, args.max_ctas = block_n, max_ctas
    if gemm_timeline is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.pk_gemm_bf16(C.byref(args), _stream()), "pk_gemm_bf16")
        e1.record()
        gemm_timeline.append((e0, e1, 2.0 * M * N * K))
        return out
    check(lib.pk_gemm_bf16(C.byref(args), _stream()), "pk_gemm_bf16")
    return out


def patchify(images: torch.Tensor, patch_size: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib_for(images)
    if images.dim() != 4 or images.shape[1] != 3 or images.shape[2] != images.shape[3] or not images.is_contiguous():
        raise ValueError(f"images must be contiguous [B,3,S,S], got {tuple(images.shape)}")
    B, _, S, _ = images.shape
    P, Kp = (S // patch_size) ** 2, 3 * patch_size * patch_size
    if out is None:
        out = torch.empty(B * P, Kp, dtype=torch.bfloat16, device=images.device)
    check(lib.pk_patchify(_ptr(images, torch.float32), _ptr(out, torch.bfloat16), B, S, patch_size, _stream()), "pk_patchify")
    return out


def fill_token_rows(x: torch.Tensor, batch: int, seq_stride: int, row_offset: int, tokens: Optional[torch.Tensor],
                    pos: Optional[torch.Tensor], scale: float = 1.0, n_tokens: Optional[int] = None) -> None:
    lib = _lib_for(x)
    dim = x.shape[-1]
    if tokens is not None:
        tokens = tokens.reshape(-1, dim)
        n_tokens = tokens.shape[0]
    if pos is not None:
        pos = pos.reshape(-1, dim)
    check(lib.pk_fill_token_rows(_ptr(x, torch.float32), batch, seq_stride, row_offset, n_tokens, dim,
                                 _ptr(tokens, torch.float32), _ptr(pos, torch.float32), float(scale), _stream()),
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
              extra_kv: Optional[torch.Tensor] = None, extra_mult: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib_for(qkv)
    D = num_heads * head_dim
    if qkv.shape[-1] != 3 * D or not qkv.is_contiguous() or not out.is_contiguous():
        raise ValueError("qkv must be contiguous [rows, 3*D] and out contiguous [rows, D]")
    a = AttentionArgs()
    a.qkv, a.out = _ptr(qkv, torch.bfloat16), _ptr(out, torch.bfloat16)
    a.batch, a.num_heads, a.head_dim = batch, num_heads, head_dim
    a.seq_len = seq_len
    a.cu_seqlens = _ptr(cu_seqlens, torch.int32)
    a.max_seq_len = max_seq_len or seq_len
    a.scale = 1.0 / math.sqrt(head_dim)
    a.key_mult = _ptr(key_mult, torch.float32)
    a.extra_kv = _ptr(extra_kv, torch.bfloat16)
    a.extra_mult = _ptr(extra_mult, torch.float32)
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


def device_flag(reset: bool = True) -> int:
    """Watchdog word of the tcgen05 kernels (0 = healthy). Synchronises the device."""
    return _lib.load().pk_device_flag(int(reset))
