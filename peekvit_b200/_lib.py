"""ctypes binding of ``libpeekvit_b200.so`` (the C ABI declared in ``include/peekvit_b200.h``).

There is no fallback: if the CUDA library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List, Optional

from . import build as _build

c_void_p, c_int, c_float, c_longlong = C.c_void_p, C.c_int, C.c_float, C.c_longlong

PK_ABI_VERSION = 2          # include/peekvit_b200.h
PK_EPI_BIAS_BF16 = 0
PK_EPI_BIAS_GELU_BF16 = 1
PK_EPI_BIAS_RESID_F32 = 2
PK_EPI_BIAS_F32 = 3
PK_OUT_BF16, PK_OUT_F16, PK_OUT_BF16X2 = 0, 1, 2


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", c_void_p), ("W", c_void_p),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("lda", c_longlong), ("ldw", c_longlong),
        ("bias", c_void_p), ("epilogue", c_int),
        ("out", c_void_p), ("ldo", c_longlong),
        ("resid", c_void_p), ("ldr", c_longlong),
        ("rowscale", c_void_p),
        ("rows_per_group", c_int), ("group_stride", c_int), ("group_offset", c_int), ("resid_is_pos", c_int),
        ("pos_offset", c_int),
        ("m_dev", c_void_p), ("row_begin_dev", c_void_p), ("out_row_index", c_void_p), ("block_n", c_int), ("max_ctas", c_int),
        ("epilogue_mode", c_int), ("cta_pair", c_int),
        ("xb_out", c_void_p), ("ldxb", c_longlong), ("row_stats", c_void_p),
        ("ln_stats", c_void_p), ("ln_c1", c_void_p), ("ln_parts", c_int), ("ln_dim", c_int), ("ln_eps", c_float),
        ("a_wrap_k", c_int), ("out_format", c_int),
        ("group_offsets", c_void_p), ("n_groups", c_int),
    ]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("qkv", c_void_p), ("out", c_void_p),
        ("batch", c_int), ("num_heads", c_int), ("head_dim", c_int),
        ("seq_len", c_int), ("cu_seqlens", c_void_p), ("max_seq_len", c_int),
        ("scale", c_float),
        ("key_mult", c_void_p), ("extra_kv", c_void_p), ("extra_mult", c_void_p),
        ("impl", c_int), ("qkv_format", c_int), ("out_format", c_int), ("route_rows", c_void_p), ("route_min_rows", c_int),
        ("total_rows", c_int),
    ]


class CompactArgs(C.Structure):
    _fields_ = [
        ("x_in", c_void_p), ("x_out", c_void_p), ("dim", c_int),
        ("cu_in", c_void_p), ("cu_out", c_void_p), ("batch", c_int),
        ("rows_in_cap", c_int),
        ("dst_local", c_void_p), ("sample_of", c_void_p),
        ("scale_in", c_void_p), ("scale_out", c_void_p),
        ("a0_in", c_void_p), ("a0_out", c_void_p),
        ("a1_in", c_void_p), ("a1_out", c_void_p),
        ("a2_in", c_void_p), ("a2_out", c_void_p),
        ("ghost", c_int),
        ("pub_tok_row", c_void_p), ("pub_mask", c_void_p), ("pub_n_img", c_int),
    ]


class ResidualGateArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("cu_in", c_void_p), ("mult_in", c_void_p),
        ("batch", c_int), ("dim", c_int), ("max_seq_len", c_int),
        ("n_special", c_int), ("budget_pos", c_int),
        ("gated", c_int),
        ("gate_w", c_void_p), ("gate_b", c_float), ("gate_temp", c_float), ("gate_bias", c_float), ("gate_type", c_int),
        ("thr_mode", c_int),
        ("bt_w", c_void_p), ("bt_b", c_float), ("thr_dev", c_void_p), ("thr_const", c_float),
        ("mask", c_void_p), ("dst_local", c_void_p), ("sample_of", c_void_p), ("new_len", c_void_p), ("mdrop", c_void_p),
    ]


class AvitArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("cu_in", c_void_p), ("batch", c_int), ("dim", c_int), ("seq_total", c_int),
        ("c", c_void_p), ("R", c_void_p), ("tokid", c_void_p),
        ("gate_scale", c_float), ("gate_center", c_float), ("eps", c_float), ("last_layer", c_int), ("early_exit", c_int),
        ("out_acc", c_void_p), ("rho", c_void_p), ("counter", c_void_p),
        ("dst_local", c_void_p), ("sample_of", c_void_p), ("new_len", c_void_p), ("n_halted", c_void_p),
    ]


# name -> (restype, argtypes); must list every symbol include/peekvit_b200.h declares
SIGNATURES: Dict[str, tuple] = {
    "pk_abi_version": (c_int, []),
    "pk_init": (c_int, [c_int]),
    "pk_last_error": (C.c_char_p, []),
    "pk_num_sms": (c_int, []),
    "pk_device_flag": (c_int, [c_int]),
    "pk_device_flag_async": (c_int, [c_void_p, c_void_p]),
    "pk_gemm_bf16": (c_int, [C.POINTER(GemmArgs), c_void_p]),
    "pk_gemm_row_stat_parts": (c_int, [c_int]),
    "pk_row_stats_cast": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "pk_patchify": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pk_patchify_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, C.POINTER(c_float), C.POINTER(c_float), c_int, c_int, c_void_p]),
    "pk_fill_token_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p]),
    "pk_layernorm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pk_attention_fwd": (c_int, [C.POINTER(AttentionArgs), c_void_p]),
    "pk_attention_trace": (c_int, [c_void_p]),
    "pk_cls_head": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "pk_argmax_count": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pk_token_norm_score": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "pk_topk_select": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "pk_gather_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "pk_exclusive_scan_i32": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "pk_compact_rows": (c_int, [C.POINTER(CompactArgs), c_void_p]),
    "pk_budget_mean_threshold": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pk_residual_gate_plan": (c_int, [C.POINTER(ResidualGateArgs), c_void_p]),
    "pk_residual_ghost": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pk_residual_publish": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pk_avit_halt_plan": (c_int, [C.POINTER(AvitArgs), c_void_p]),
    "pk_scatter_add_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pk_cls_features": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "pk_expert_onehot": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pk_split3_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pk_patchify_split3": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "pk_split2_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pk_patchify_split2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "pk_attention_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pk_noise_snr": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "pk_cast_f32_bf16": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "pk_gelu_bf16": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "pk_gelu_bwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "pk_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "pk_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "pk_softmax_xent": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pk_head_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pk_scatter_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "pk_residual_gate_train_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_float, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pk_residual_gate_train_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                           c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pk_layernorm_bwd_gated": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                       c_int, c_void_p]),
    "pk_cast_rows_f32_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pk_rowdot": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p]),
    "pk_sum_token_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pk_row_scale_add": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "pk_zero_token_rows": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "pk_moe_route": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
}

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "peekvit_b200.h")


def header_symbols() -> List[str]:
    """Function names declared in include/peekvit_b200.h."""
    with open(HEADER) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", src)))


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load (building first if the library is absent) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PEEKVIT_B200_LIB") or _build.LIB_PATH      # override: A/B runs of two builds on the same box
    if path == _build.LIB_PATH:
        # never run a binary that was built from other sources than the ones in the tree (content hash, not file times)
        if os.environ.get("PEEKVIT_B200_REBUILD") == "1" or _build.is_stale():
            if not _build.have_nvcc():
                raise RuntimeError(f"{path} is missing or was built from different sources and nvcc is not available to rebuild it")
            path = _build.build(force=True)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.pk_abi_version() != PK_ABI_VERSION:
        raise RuntimeError(f"libpeekvit_b200 ABI {lib.pk_abi_version()} != {PK_ABI_VERSION}; rebuild with `python -m peekvit_b200.build --force`")
    _lib = lib
    return lib


class PkError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().pk_last_error()
        raise PkError(f"{what or 'peekvit_b200'} failed (status {rc}): {msg.decode() if msg else ''}")


_initialised_devices = set()


def init(device_index: int) -> C.CDLL:
    """Create the library context of ``device_index`` (must be a B200 / sm_100a); one context per device of the process."""
    lib = load()
    if device_index not in _initialised_devices:
        check(lib.pk_init(device_index), "pk_init")
        _initialised_devices.add(device_index)
    return lib
