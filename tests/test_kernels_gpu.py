"""GPU parity tests of every C-ABI kernel against the torch fp32 expression at the cited
reference call site.  All calls go through ``peekvit_b200.ops`` -> ctypes -> libpeekvit_b200.so.

Tolerances (written here on purpose): fp32 outputs 2e-5 relative to max|ref| (fp32 accumulate of
bf16 operands is compared with an fp32 matmul of the *same* bf16-rounded operands); bf16 outputs
1e-2 relative (north star), typically 3e-3 = one bf16 rounding; index outputs bit-exact.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
TOL_F32 = 2e-5
TOL_BF16 = 1e-2


@pytest.fixture(scope="module")
def ops():
    from peekvit_b200 import ops as _ops
    torch.backends.cuda.matmul.allow_tf32 = False
    return _ops


def rel_err(got, ref):
    got, ref = got.float(), ref.float()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def _operands(M, N, K, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    a = (torch.randn(M, K, device=DEV, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    return a, w, bias


# ------------------------------------------------------------------ GEMM (K1,K3,K5,K6,K7)
@pytest.mark.parametrize("shape", [(1, 64, 64), (127, 192, 64), (128, 128, 128), (129, 256, 64), (300, 768, 768),
                                   (1000, 2304, 768), (1576, 3072, 768), (257, 768, 3072), (64, 128, 192),
                                   (394, 1152, 384), (785 * 2, 768, 256), (33, 576, 192), (513, 1000, 768)])
@pytest.mark.parametrize("block_n", [0, 128, 192, 256])
def test_gemm_f32(ops, shape, block_n):
    from peekvit_b200._lib import PK_EPI_BIAS_F32
    M, N, K = shape
    a, w, bias = _operands(M, N, K)
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(a, w, bias, out, PK_EPI_BIAS_F32, block_n=block_n)
    assert ops.device_flag() == 0
    assert rel_err(out, a.float() @ w.float().t() + bias) < TOL_F32


@pytest.mark.parametrize("shape", [(127, 192, 64), (300, 768, 768), (1000, 2304, 768), (513, 1000, 768)])
def test_gemm_simt_epilogue_path(ops, shape):
    """epilogue_mode=2 forces the per-row-predicated SIMT epilogue that remapped / segmented rows use."""
    from peekvit_b200._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_F32, PK_EPI_BIAS_GELU_BF16, PK_EPI_BIAS_RESID_F32
    M, N, K = shape
    a, w, bias = _operands(M, N, K, seed=9)
    acc = a.float() @ w.float().t() + bias
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(a, w, bias, out, PK_EPI_BIAS_F32, epilogue_mode=2)
    assert rel_err(out, acc) < TOL_F32
    x = torch.randn(M, N, device=DEV)
    x0 = x.clone()
    ops.gemm(a, w, bias, x, PK_EPI_BIAS_RESID_F32, resid=x, epilogue_mode=2)
    assert rel_err(x, acc + x0) < TOL_F32
    ob = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, bias, ob, PK_EPI_BIAS_GELU_BF16, epilogue_mode=2)
    assert rel_err(ob, torch.nn.functional.gelu(acc)) < TOL_BF16
    ops.gemm(a, w, bias, ob, PK_EPI_BIAS_BF16, epilogue_mode=2)
    assert rel_err(ob, acc) < TOL_BF16
    assert ops.device_flag() == 0


def test_gemm_epilogues(ops):
    from peekvit_b200._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16, PK_EPI_BIAS_RESID_F32
    M, N, K = 1000, 768, 768
    a, w, bias = _operands(M, N, K, seed=1)
    acc = a.float() @ w.float().t() + bias
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, bias, out, PK_EPI_BIAS_BF16)
    assert rel_err(out, acc) < TOL_BF16
    ops.gemm(a, w, bias, out, PK_EPI_BIAS_GELU_BF16)
    assert rel_err(out, torch.nn.functional.gelu(acc)) < TOL_BF16      # exact (erf) GELU, blocks.py:82
    x = torch.randn(M, N, device=DEV)
    x0, rs = x.clone(), torch.rand(M, device=DEV)
    ops.gemm(a, w, bias, x, PK_EPI_BIAS_RESID_F32, resid=x, rowscale=rs)   # in place on the residual stream
    assert rel_err(x, rs[:, None] * acc + x0) < TOL_F32
    ops.gemm(a, w, None, x, PK_EPI_BIAS_RESID_F32, resid=x0)               # no bias, out-of-place residual
    assert rel_err(x, a.float() @ w.float().t() + x0) < TOL_F32
    assert ops.device_flag() == 0


def test_gelu_epilogue_accuracy_fp32_grid(ops):
    """The in-kernel erf (A&S 7.1.26) against torch's exact GELU over a dense grid, before bf16
    rounding matters: identity weight trick (acc = x exactly representable in bf16)."""
    from peekvit_b200._lib import PK_EPI_BIAS_GELU_BF16
    K = 64
    xs = torch.linspace(-8, 8, 128 * 64, device=DEV).to(torch.bfloat16).view(128, 64)
    eye = torch.eye(K, device=DEV, dtype=torch.bfloat16)
    out = torch.empty(128, 64, device=DEV, dtype=torch.bfloat16)
    ops.gemm(xs, eye, None, out, PK_EPI_BIAS_GELU_BF16)
    ref = torch.nn.functional.gelu(xs.float())
    # within one bf16 ulp of torch's exact GELU everywhere on the grid
    assert ((out.float() - ref).abs() <= ref.abs() * 2.0 ** -8 + 2e-6).all()


def test_gemm_patch_remap_and_device_m(ops):
    from peekvit_b200._lib import PK_EPI_BIAS_F32, PK_EPI_BIAS_RESID_F32
    G, P, seq, off, D, Kp = 5, 60, 66, 3, 256, 192
    a, w, bias = _operands(G * P, D, Kp, seed=2)
    pos = torch.randn(seq, D, device=DEV)
    x = torch.zeros(G * seq, D, device=DEV)
    ops.gemm(a, w, bias, x, PK_EPI_BIAS_RESID_F32, resid=pos, rows_per_group=P, group_stride=seq, group_offset=off,
             resid_is_pos=True)
    ref = torch.zeros(G, seq, D, device=DEV)
    ref[:, off:off + P] = (a.float() @ w.float().t() + bias).view(G, P, D) + pos[off:off + P]
    assert rel_err(x, ref.view(G * seq, D)) < TOL_F32
    assert (x.view(G, seq, D)[:, :off] == 0).all() and (x.view(G, seq, D)[:, off + P:] == 0).all()
    M, N, K = 1000, 384, 384
    a, w, _ = _operands(M, N, K, seed=3)
    out = torch.zeros(M, N, device=DEV)
    for m in (0, 1, 333, 1000):
        out.zero_()
        ops.gemm(a, w, None, out, PK_EPI_BIAS_F32, m_dev=torch.tensor([m], device=DEV, dtype=torch.int32))
        ref = a.float() @ w.float().t()
        if m:
            assert rel_err(out[:m], ref[:m]) < TOL_F32
        assert (out[m:] == 0).all()
    assert ops.device_flag() == 0


def test_gemm_rejects_bad_arguments(ops):
    from peekvit_b200._lib import PK_EPI_BIAS_F32, PkError
    a, w, bias = _operands(64, 64, 96)          # K not a multiple of 64
    with pytest.raises(PkError):
        ops.gemm(a, w, bias, torch.empty(64, 64, device=DEV), PK_EPI_BIAS_F32)
    with pytest.raises(TypeError):
        ops.gemm(a.float(), w, bias, torch.empty(64, 64, device=DEV), PK_EPI_BIAS_F32)
    with pytest.raises(RuntimeError):
        ops.layernorm(torch.randn(4, 64), torch.ones(64), torch.zeros(64), 1e-5)   # CPU tensor: no CPU path


def test_gemm_linearity_full_size(ops):
    """Size-independent property at the BASELINE config-B layer shape (M = 197*256 rows):
    GEMM(a, w1 + w2) == GEMM(a, w1) + GEMM(a, w2) with exactly representable operands."""
    from peekvit_b200._lib import PK_EPI_BIAS_F32
    M, N, K = 197 * 256, 768, 768
    g = torch.Generator(device=DEV).manual_seed(5)
    a = torch.randint(-4, 5, (M, K), device=DEV, generator=g).to(torch.bfloat16)
    w1 = torch.randint(-4, 5, (N, K), device=DEV, generator=g).to(torch.bfloat16)
    w2 = torch.randint(-4, 5, (N, K), device=DEV, generator=g).to(torch.bfloat16)
    o1, o2, o3 = (torch.empty(M, N, device=DEV) for _ in range(3))
    ops.gemm(a, w1, None, o1, PK_EPI_BIAS_F32)
    ops.gemm(a, w2, None, o2, PK_EPI_BIAS_F32)
    ops.gemm(a, (w1.float() + w2.float()).to(torch.bfloat16), None, o3, PK_EPI_BIAS_F32)
    assert torch.equal(o1 + o2, o3)             # small integers: every partial sum is exact in fp32
    assert ops.device_flag() == 0


# ------------------------------------------------------------------ LayerNorm (K2)
@pytest.mark.parametrize("D", [64, 128, 192, 256, 384, 768, 1024])
@pytest.mark.parametrize("eps", [1e-5, 1e-6])
def test_layernorm(ops, D, eps):
    g = torch.Generator(device=DEV).manual_seed(D)
    x = torch.randn(1237, D, device=DEV, generator=g) * 2 + 0.3
    gamma, beta = torch.randn(D, device=DEV, generator=g), torch.randn(D, device=DEV, generator=g)
    y = ops.layernorm(x, gamma, beta, eps)
    assert rel_err(y, torch.nn.functional.layer_norm(x, (D,), gamma, beta, eps)) < TOL_BF16


def test_layernorm_rowscale_gather_devrows(ops):
    x = torch.randn(1000, 384, device=DEV)
    gamma, beta = torch.randn(384, device=DEV), torch.randn(384, device=DEV)
    rs = torch.rand(500, device=DEV)
    idx = torch.randperm(1000, device=DEV)[:500].to(torch.int32)
    y = ops.layernorm(x, gamma, beta, 1e-6, rowscale=rs, row_index=idx)
    ref = rs[:, None] * torch.nn.functional.layer_norm(x[idx.long()], (384,), gamma, beta, 1e-6)
    assert rel_err(y, ref) < TOL_BF16
    y = torch.zeros(1000, 384, device=DEV, dtype=torch.bfloat16)
    ops.layernorm(x, gamma, beta, 1e-5, y, rows_dev=torch.tensor([123], device=DEV, dtype=torch.int32))
    assert (y[123:] == 0).all() and y[:123].abs().sum() > 0
    # zero rows (a masked ResidualViT token): LN(0) = beta
    y = ops.layernorm(torch.zeros(3, 384, device=DEV), gamma, beta, 1e-6)
    assert rel_err(y, beta.expand(3, -1)) < TOL_BF16


# ------------------------------------------------------------------ attention (K4)
def ref_attention(qkv, B, H, dh, lens, key_mult=None, extra_kv=None, extra_mult=None):
    """softmax(q k^T / sqrt(dh) [+ log mult]) v per sample and head, fp32 (torch functional MHA math)."""
    D = H * dh
    out = torch.zeros(qkv.shape[0], D, device=qkv.device)
    start = 0
    for b in range(B):
        n = lens[b]
        blk = qkv[start:start + n].float()
        q, k, v = (blk[:, i * D:(i + 1) * D].view(n, H, dh).transpose(0, 1) for i in range(3))
        bias = torch.zeros(n, device=qkv.device) if key_mult is None else key_mult[start:start + n].log()
        if extra_kv is not None and extra_mult[b] > 0:
            k = torch.cat([k, extra_kv[:D].float().view(H, 1, dh)], 1)
            v = torch.cat([v, extra_kv[D:].float().view(H, 1, dh)], 1)
            bias = torch.cat([bias, extra_mult[b:b + 1].log()])
        o = torch.softmax((q @ k.transpose(1, 2)) / math.sqrt(dh) + bias, -1) @ v
        out[start:start + n] = o.transpose(0, 1).reshape(n, D)
        start += n
    return out


@pytest.mark.parametrize("cfg", [(3, 2, 64, 17), (2, 12, 64, 197), (2, 8, 32, 785), (4, 6, 64, 64), (1, 3, 64, 1), (2, 8, 48, 197), (3, 2, 48, 70),
                                 (2, 2, 32, 65), (3, 6, 64, 198), (1, 12, 64, 257)])
def test_attention_dense(ops, cfg):
    B, H, dh, N = cfg
    D = H * dh
    qkv = torch.randn(B * N, 3 * D, device=DEV).to(torch.bfloat16)
    out = torch.zeros(B * N, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, H, dh, seq_len=N)
    assert rel_err(out, ref_attention(qkv, B, H, dh, [N] * B)) < TOL_BF16


def test_attention_ragged_multiplicity_virtual_key(ops):
    B, H, dh = 6, 6, 64
    D = H * dh
    lens = [3, 70, 198, 1, 129, 64]
    cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), device=DEV, dtype=torch.int32)
    rows = sum(lens)
    qkv = torch.randn(rows, 3 * D, device=DEV).to(torch.bfloat16)
    out = torch.zeros(rows, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, H, dh, cu_seqlens=cu, max_seq_len=max(lens))
    assert rel_err(out, ref_attention(qkv, B, H, dh, lens)) < TOL_BF16
    km = torch.randint(1, 40, (rows,), device=DEV).float()
    km[5] = 0.0                                          # multiplicity 0 = dead row: never attended to
    ekv = (torch.randn(2 * D, device=DEV) * 0.5).to(torch.bfloat16)
    em = torch.tensor([0.0, 5.0, 100.0, 7.0, 0.0, 1.0], device=DEV)   # sample 5: 64 keys + virtual key -> second tile
    ops.attention(qkv, out, B, H, dh, cu_seqlens=cu, max_seq_len=max(lens), key_mult=km, extra_kv=ekv, extra_mult=em)
    assert rel_err(out, ref_attention(qkv, B, H, dh, lens, km, ekv, em)) < TOL_BF16


def test_attention_equals_dense_reference_with_zero_tokens(ops):
    """The compaction identity the sparse models rely on: M zeroed tokens
    inside a dense sequence == one virtual bias key with multiplicity M."""
    H, dh, n_live, n_dead = 6, 64, 50, 30
    D = H * dh
    bias_kv = (torch.randn(2 * D, device=DEV) * 0.3).to(torch.bfloat16)
    live = torch.randn(n_live, 3 * D, device=DEV).to(torch.bfloat16)
    dead = torch.cat([torch.zeros(D, device=DEV, dtype=torch.bfloat16), bias_kv]).expand(n_dead, -1)   # k=b_k, v=b_v
    dense = torch.cat([live, dead]).contiguous()
    out_dense = torch.zeros(n_live + n_dead, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(dense, out_dense, 1, H, dh, seq_len=n_live + n_dead)
    out_sparse = torch.zeros(n_live, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(live.contiguous(), out_sparse, 1, H, dh, seq_len=n_live, extra_kv=bias_kv,
                  extra_mult=torch.tensor([float(n_dead)], device=DEV))
    assert rel_err(out_sparse, out_dense[:n_live]) < TOL_BF16


# ------------------------------------------------------------------ patchify / token rows / head (K1,K8)
@pytest.mark.parametrize("S,p", [(64, 8), (224, 16), (32, 8), (224, 8), (48, 16)])
def test_patchify(ops, S, p):
    img = torch.randn(3, 3, S, S, device=DEV)
    got = ops.patchify(img, p)
    ref = torch.nn.functional.unfold(img, kernel_size=p, stride=p).transpose(1, 2).reshape(3 * (S // p) ** 2, 3 * p * p)
    assert torch.equal(got, ref.to(torch.bfloat16))        # (c,i,j) K order == conv_proj.weight.reshape(D,-1)


@pytest.mark.parametrize("u8", [False, True])
def test_patchify_token_row_layout(ops, u8):
    """rows_per_sample / row_offset: patch q of sample b lands on row b*rows_per_sample + row_offset + q, other rows untouched."""
    B, S, p, seq, off = 3, 64, 8, 67, 2
    P = (S // p) ** 2
    if u8:
        img = torch.randint(0, 256, (B, S, S, 3), device=DEV, dtype=torch.uint8)
        dense = ops.patchify_u8(img, p)
        out = torch.full((B * seq, 3 * p * p), 7.0, device=DEV, dtype=torch.bfloat16)
        ops.patchify_u8(img, p, out, rows_per_sample=seq, row_offset=off)
    else:
        img = torch.randn(B, 3, S, S, device=DEV)
        dense = ops.patchify(img, p)
        out = torch.full((B * seq, 3 * p * p), 7.0, device=DEV, dtype=torch.bfloat16)
        ops.patchify(img, p, out, rows_per_sample=seq, row_offset=off)
    o3 = out.view(B, seq, -1)
    assert torch.equal(o3[:, off:off + P], dense.view(B, P, -1))
    assert (o3[:, :off] == 7).all() and (o3[:, off + P:] == 7).all()


def test_patch_embed_token_rows_matches_remap(ops):
    """Token-row operand + rowscale-masked staged residual (CTA-pair kernel, with the LayerNorm producer outputs) gives the
    same residual stream as the densely packed operand + row-remap epilogue, bit for bit."""
    from peekvit_b200._lib import PK_EPI_BIAS_RESID_F32
    B, S, p, D, T = 6, 64, 8, 128, 1
    P, seq = 64, 65
    img = torch.randn(B, 3, S, S, device=DEV)
    w = (torch.randn(D, 3 * p * p, device=DEV) / math.sqrt(3 * p * p)).to(torch.bfloat16)
    bias, pos, cls = torch.randn(D, device=DEV) * 0.1, torch.randn(seq, D, device=DEV) * 0.02, torch.randn(1, D, device=DEV)
    x_ref = torch.zeros(B * seq, D, device=DEV)
    ops.gemm(ops.patchify(img, p), w, bias, x_ref, PK_EPI_BIAS_RESID_F32, resid=pos, rows_per_group=P, group_stride=seq,
             group_offset=T, resid_is_pos=True)
    ops.fill_token_rows(x_ref, B, seq, 0, cls, pos)
    a = torch.zeros(B * seq, 3 * p * p, device=DEV, dtype=torch.bfloat16)
    ops.patchify(img, p, a, rows_per_sample=seq, row_offset=T)
    x_init = torch.zeros(B * seq, D, device=DEV)
    ops.fill_token_rows(x_init, B, seq, 0, cls, pos)
    ops.fill_token_rows(x_init, B, seq, T, None, pos, n_tokens=P, scale=0.0)
    rs = torch.zeros(B, seq, device=DEV)
    rs[:, T:] = 1
    x = torch.full((B * seq, D), float("nan"), device=DEV)
    xb = torch.zeros(B * seq, D, device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(B * seq, ops.gemm_row_stat_parts(D), 2, device=DEV)
    ops.gemm(a, w, bias, x, PK_EPI_BIAS_RESID_F32, resid=x_init, rowscale=rs.reshape(-1), xb_out=xb, row_stats=stats, cta_pair=2)
    assert torch.equal(x, x_ref)
    assert torch.equal(xb, x_ref.to(torch.bfloat16))
    s = stats.sum(1)
    assert rel_err(s[:, 0], x_ref.sum(1)) < 1e-5 and rel_err(s[:, 1], (x_ref * x_ref).sum(1)) < 1e-5


def test_patch_embed_matches_conv2d(ops):
    """patchify + GEMM(+bias+pos) == Conv2d(k=s=p) + flatten + transpose + pos (vit.py:212-220,:92)."""
    from peekvit_b200._lib import PK_EPI_BIAS_RESID_F32
    B, S, p, D = 4, 64, 8, 128
    img = torch.randn(B, 3, S, S, device=DEV)
    wt = (torch.randn(D, 3, p, p, device=DEV) / math.sqrt(3 * p * p))
    bias, pos = torch.randn(D, device=DEV) * 0.1, torch.randn(65, D, device=DEV) * 0.02
    x = torch.zeros(B * 65, D, device=DEV)
    ops.gemm(ops.patchify(img, p), wt.reshape(D, -1).to(torch.bfloat16).contiguous(), bias, x, PK_EPI_BIAS_RESID_F32, resid=pos,
             rows_per_group=64, group_stride=65, group_offset=1, resid_is_pos=True)
    ref = torch.nn.functional.conv2d(img.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), bias, stride=p)
    ref = ref.reshape(B, D, 64).permute(0, 2, 1) + pos[1:]
    assert rel_err(x.view(B, 65, D)[:, 1:], ref) < TOL_F32


def test_fill_token_rows(ops):
    B, seq, T, D = 3, 20, 2, 256
    x = torch.zeros(B * seq, D, device=DEV)
    tok, pos = torch.randn(T, D, device=DEV), torch.randn(seq, D, device=DEV)
    ops.fill_token_rows(x, B, seq, 1, tok, pos, scale=0.4)
    ref = torch.zeros(B, seq, D, device=DEV)
    ref[:, 1:1 + T] = 0.4 * tok + pos[1:1 + T]
    assert rel_err(x, ref.view(B * seq, D)) < 1e-6
    x.zero_()
    ops.fill_token_rows(x, B, seq, seq - 1, None, None, scale=0.7, n_tokens=1)
    xv = x.view(B, seq, D)
    assert (xv[:, -1] == 0.7).all() and (xv[:, :-1] == 0).all()


@pytest.mark.parametrize("cfg", [(5, 128, 10, 1), (19, 768, 1000, 1), (4, 384, 7, 2), (1, 64, 10, 1), (9, 256, 10, 1)])
def test_cls_head(ops, cfg):
    B, D, C, T = cfg
    seq = 11
    x = torch.randn(B * seq, D, device=DEV)
    gamma, beta = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    hw, hb = torch.randn(C, D, device=DEV) / math.sqrt(D), torch.randn(C, device=DEV)
    got = ops.cls_head(x, B, seq, T, gamma, beta, 1e-5, hw, hb)
    f = torch.nn.functional.layer_norm(x.view(B, seq, D)[:, :T], (D,), gamma, beta, 1e-5).sum(1)    # class-token SUM
    assert rel_err(got, f @ hw.t() + hb) < TOL_F32
    cu = (torch.arange(B + 1, device=DEV) * seq).to(torch.int32)
    got2 = ops.cls_head(x, B, 0, T, gamma, beta, 1e-5, hw, hb, cu_seqlens=cu)
    assert torch.equal(got, got2)


# ------------------------------------------------------------------ RankViT kernels (K9,K10,K11)
def test_token_norm_score(ops):
    B, seq, D = 7, 197, 768
    x = torch.randn(B * seq, D, device=DEV)
    sc = ops.token_norm_score(x, B, seq)
    assert rel_err(sc, torch.norm(x.view(B, seq, D)[:, 1:], dim=-1)) < 1e-6


@pytest.mark.parametrize("n,k", [(196, 1), (196, 98), (196, 196), (784, 392), (49, 25), (1, 1), (4096, 1000)])
def test_topk_bit_exact(ops, n, k):
    sc = torch.randn(5, n, device=DEV)
    kept = ops.topk_select(sc, k)
    assert torch.equal(kept.long(), torch.argsort(sc, dim=-1, descending=True, stable=True)[:, :k])


def test_topk_ties_lowest_index(ops):
    adv = torch.tensor([[1., 3, 3, 0, 3, 1, -0., 0.], [2.] * 8, [0., -0., 0., 1e-45, -0., 5, 5, 5]], device=DEV)
    for k in (1, 3, 8):
        kept = ops.topk_select(adv, k)
        assert torch.equal(kept.cpu().long(), torch.argsort(adv.cpu(), dim=-1, descending=True, stable=True)[:, :k])
    assert ops.topk_select(adv, 3).cpu().tolist() == [[1, 2, 4], [0, 1, 2], [5, 6, 7]]
    big = torch.randn(3, 4096, device=DEV).round(decimals=1)       # ~100 distinct values: ties everywhere
    assert torch.equal(ops.topk_select(big, 1000).long(), torch.argsort(big, dim=-1, descending=True, stable=True)[:, :1000])


def test_gather_rows_and_sort_and_drop(ops):
    """score -> select -> compact == reference sort_and_drop (rankvit.py:55-77) on the same input."""
    B, seq, D, budget = 6, 197, 768, 0.5
    x = torch.randn(B * seq, D, device=DEV)
    k = math.ceil((seq - 1) * budget)
    kept = ops.topk_select(ops.token_norm_score(x, B, seq), k)
    y = ops.gather_rows(x, kept, B, seq).view(B, k + 1, D)
    xr = x.view(B, seq, D)
    cls, tok = xr[:, :1], xr[:, 1:]
    idx = torch.argsort(torch.norm(tok, dim=-1), dim=-1, descending=True, stable=True).unsqueeze(-1)
    ref = torch.cat([cls, torch.gather(tok, 1, idx.expand(-1, -1, D))[:, :k]], 1)
    assert torch.equal(y, ref)


# ------------------------------------------------------------------ CTA-pair GEMM (cta_group::2) and the single-CTA kernel, forced
@pytest.mark.parametrize("shape", [(256, 256, 64), (300, 768, 768), (1000, 2304, 768), (257, 768, 3072), (777, 1152, 384),
                                   (600, 1000, 768), (197 * 40, 3072, 768)])
@pytest.mark.parametrize("cta_pair", [1, 2])
def test_gemm_both_kernels_all_epilogues(ops, shape, cta_pair):
    """cta_pair=2 forces the 256 x BN tcgen05 cta_group::2 kernel, 1 the 128 x BN single-CTA kernel; same results.
    The in-place residual epilogue of the pair kernel adds through a TMA reduction at L2 (x += ...)."""
    from peekvit_b200._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_F32, PK_EPI_BIAS_GELU_BF16, PK_EPI_BIAS_RESID_F32
    M, N, K = shape
    a, w, bias = _operands(M, N, K, seed=3)
    acc = a.float() @ w.float().t() + bias
    for bn in (0, 128):
        out = torch.full((M, N), float("nan"), device=DEV)
        ops.gemm(a, w, bias, out, PK_EPI_BIAS_F32, block_n=bn, cta_pair=cta_pair)
        assert rel_err(out, acc) < TOL_F32
        outb = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
        ops.gemm(a, w, bias, outb, PK_EPI_BIAS_BF16, block_n=bn, cta_pair=cta_pair)
        assert rel_err(outb, acc) < TOL_BF16
        ops.gemm(a, w, bias, outb, PK_EPI_BIAS_GELU_BF16, block_n=bn, cta_pair=cta_pair)
        assert rel_err(outb, torch.nn.functional.gelu(acc)) < TOL_BF16
        x = torch.randn(M, N, device=DEV)
        rs = torch.rand(M, device=DEV)
        x0 = x.clone()
        ops.gemm(a, w, bias, x, PK_EPI_BIAS_RESID_F32, resid=x, rowscale=rs, block_n=bn, cta_pair=cta_pair)      # in place
        assert rel_err(x, rs[:, None] * acc + x0) < TOL_F32
        y = torch.empty(M, N, device=DEV)
        ops.gemm(a, w, bias, y, PK_EPI_BIAS_RESID_F32, resid=x0, block_n=bn, cta_pair=cta_pair)                  # out of place
        assert rel_err(y, acc + x0) < TOL_F32
    assert ops.device_flag() == 0


@pytest.mark.parametrize("m_dev", [0, 1, 255, 256, 333, 1000])
def test_pair_gemm_device_side_row_count(ops, m_dev):
    """Ragged batches: rows >= *m_dev must stay untouched (bit-exact), rows below match."""
    from peekvit_b200._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_F32, PK_EPI_BIAS_RESID_F32
    M, N, K = 1000, 768, 384
    a, w, _ = _operands(M, N, K, seed=5)
    ref = a.float() @ w.float().t()
    md = torch.tensor([m_dev], device=DEV, dtype=torch.int32)
    out = torch.full((M, N), 7.0, device=DEV)
    ops.gemm(a, w, None, out, PK_EPI_BIAS_F32, m_dev=md, cta_pair=2)
    outb = torch.full((M, N), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, None, outb, PK_EPI_BIAS_BF16, m_dev=md, cta_pair=2)
    x = torch.ones(M, N, device=DEV)
    ops.gemm(a, w, None, x, PK_EPI_BIAS_RESID_F32, resid=x, m_dev=md, cta_pair=2)
    assert bool((out[m_dev:] == 7).all()) and bool((outb[m_dev:] == 7).all()) and bool((x[m_dev:] == 1).all())
    if m_dev:
        assert rel_err(out[:m_dev], ref[:m_dev]) < TOL_F32
        assert rel_err(outb[:m_dev], ref[:m_dev]) < TOL_BF16
        assert rel_err(x[:m_dev], ref[:m_dev] + 1) < TOL_F32
    assert ops.device_flag() == 0


# ------------------------------------------------------------------ tcgen05 / TMEM attention (uniform 128 < n <= 256, head_dim 64)
@pytest.mark.parametrize("cfg", [(1, 1, 197), (2, 12, 197), (3, 6, 198), (2, 3, 129), (2, 2, 256), (2, 2, 144), (5, 4, 145),
                                 (3, 2, 160), (2, 2, 176), (2, 2, 192), (2, 2, 209), (2, 2, 225), (2, 2, 241),
                                 (3, 2, 17), (2, 3, 33), (2, 2, 50), (2, 2, 64), (3, 2, 80), (2, 12, 99), (2, 2, 113), (2, 2, 128)])
def test_attention_tcgen05(ops, cfg):
    """impl=2 forces the tcgen05/TMEM kernel (every padded length 32..256; n <= 128 runs with the second query tile
    empty); impl=1 the general mma.sync kernel."""
    B, H, N = cfg
    dh, D = 64, H * 64
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + N)
    qkv = (torch.randn(B * N, 3 * D, device=DEV, generator=g) * 1.5).to(torch.bfloat16)
    out = torch.full((B * N, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, H, dh, seq_len=N, impl=2)
    assert ops.device_flag() == 0
    blk = qkv.float().view(B, N, 3, H, dh)
    q, k, v = blk[:, :, 0].transpose(1, 2), blk[:, :, 1].transpose(1, 2), blk[:, :, 2].transpose(1, 2)
    ref = (torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), -1) @ v).transpose(1, 2).reshape(B * N, D)
    assert rel_err(out, ref) < TOL_BF16
    out1 = torch.zeros_like(out)
    ops.attention(qkv, out1, B, H, dh, seq_len=N, impl=1)
    assert rel_err(out1, ref) < TOL_BF16


def test_attention_tcgen05_full_size_rows_sum_property(ops):
    """BASELINE config size (256 images x 12 heads x 197 tokens): with V = ones every output element is exactly 1
    (softmax rows sum to one), whatever Q and K are: a size-independent check of the P V path and the row sums."""
    B, H, N, dh = 256, 12, 197, 64
    D = H * dh
    qkv = torch.randn(B * N, 3 * D, device=DEV).to(torch.bfloat16)
    qkv[:, 2 * D:] = 1.0
    out = torch.zeros(B * N, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, H, dh, seq_len=N, impl=2)
    assert ops.device_flag() == 0
    assert float((out.float() - 1.0).abs().max()) < 1e-2


def test_argmax_count(ops):
    """pk_argmax_count == torch.argmax (first maximum on ties) and the accuracy counters accumulate."""
    g = torch.Generator(device=DEV).manual_seed(4)
    logits = torch.randn(1000, 1000, device=DEV, generator=g).round(decimals=1)      # plenty of exact ties
    labels = torch.randint(0, 1000, (1000,), device=DEV, generator=g)
    ref = logits.argmax(1)
    labels[::3] = ref[::3]
    pred = ops.argmax_count(logits)
    assert torch.equal(pred.long(), ref)
    counts = torch.zeros(2, dtype=torch.int64, device=DEV)
    ops.argmax_count(logits, labels, counts)
    ops.argmax_count(logits[:10], labels[:10], counts)
    exp = int((ref == labels).sum()) + int((ref[:10] == labels[:10]).sum())
    assert counts.tolist() == [exp, 1010]
    small = torch.tensor([[1.0, 3.0, 3.0], [float("nan"), 0.0, -1.0]], device=DEV)
    assert ops.argmax_count(small).tolist() == [1, 1]


@pytest.mark.parametrize("shape", [(1000, 768, 3072), (788, 384, 1536), (50432 // 8, 768, 2304), (300, 256, 768)])
def test_gemm_layernorm_fused_chain(ops, shape):
    """out-proj(+residual) -> LayerNorm -> fc1(+GELU) as two GEMMs with the LayerNorm folded across them
    (producer epilogue: bf16 copy + row statistics; consumer epilogue: rstd / mean / folded gain), against the
    fp32 expression x1 = x0 + att Wo^T + bo; y = gelu(LN(x1) W1^T + b1)  (vit.py:48-55)."""
    from peekvit_b200._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16, PK_EPI_BIAS_RESID_F32
    M, D, F = shape
    g = torch.Generator(device=DEV).manual_seed(M + D)
    att = (torch.randn(M, D, device=DEV, generator=g) * 0.5).to(torch.bfloat16)
    wo = (torch.randn(D, D, device=DEV, generator=g) / math.sqrt(D)).to(torch.bfloat16)
    bo = torch.randn(D, device=DEV, generator=g) * 0.1
    x0 = torch.randn(M, D, device=DEV, generator=g) * 2 + 0.3            # non-zero mean: the fold must cancel it
    gamma = 1 + 0.2 * torch.randn(D, device=DEV, generator=g)
    beta = 0.1 * torch.randn(D, device=DEV, generator=g)
    w1 = torch.randn(F, D, device=DEV, generator=g) / math.sqrt(D)
    b1 = torch.randn(F, device=DEV, generator=g) * 0.1
    eps = 1e-5
    x1_ref = x0 + att.float() @ wo.float().t() + bo
    ln_ref = torch.nn.functional.layer_norm(x1_ref, (D,), gamma, beta, eps)
    lin_ref = ln_ref @ w1.t() + b1
    # ours
    P = ops.gemm_row_stat_parts(D)
    x = x0.clone()
    xb = torch.full((M, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    stats = torch.full((M, P, 2), float("nan"), device=DEV)
    ops.gemm(att, wo, bo, x, PK_EPI_BIAS_RESID_F32, resid=x, xb_out=xb, row_stats=stats)
    assert ops.device_flag() == 0
    assert rel_err(x, x1_ref) < TOL_F32
    assert rel_err(xb, x1_ref) < TOL_BF16
    s = stats.sum(1)
    assert rel_err(s[:, 0], x1_ref.sum(1)) < 1e-4 and rel_err(s[:, 1], (x1_ref * x1_ref).sum(1)) < 1e-4
    w1f = (w1 * gamma[None, :]).to(torch.bfloat16)
    c1 = w1f.float().sum(1)
    c2 = b1 + w1 @ beta
    out = torch.empty(M, F, device=DEV, dtype=torch.bfloat16)
    ops.gemm(xb, w1f, c2, out, PK_EPI_BIAS_BF16, ln_stats=stats, ln_c1=c1, ln_dim=D, ln_eps=eps)
    assert rel_err(out, lin_ref) < TOL_BF16
    ops.gemm(xb, w1f, c2, out, PK_EPI_BIAS_GELU_BF16, ln_stats=stats, ln_c1=c1, ln_dim=D, ln_eps=eps)
    assert rel_err(out, torch.nn.functional.gelu(lin_ref)) < TOL_BF16
    # statistics of rows that do not come from a producer GEMM
    xb2 = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    st2 = torch.full((M, P, 2), float("nan"), device=DEV)
    ops.row_stats_cast(x1_ref.contiguous(), xb2, st2)
    assert rel_err(st2.sum(1)[:, 0], x1_ref.sum(1)) < 1e-5 and torch.equal(xb2, x1_ref.to(torch.bfloat16))
    assert ops.device_flag() == 0


# ------------------------------------------------------------------ fp32-accurate mode (csrc/pk_exact.cu)
TOL_EXACT = 1e-5          # north star: fp32 mode within 1e-5


def _split3_ref(x):
    h = x.to(torch.bfloat16)
    r = x - h.float()
    m = r.to(torch.bfloat16)
    l = (r - m.float()).to(torch.bfloat16)
    return h, m, l


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_split3_rows(ops, mode):
    """[m|l|h|m|h|h] segments, bit-exact against the same split written in torch; the three terms reproduce the fp32 value."""
    rows, K = 37, 384
    x = torch.randn(rows, K, device=DEV) * 3
    g, b = torch.randn(K, device=DEV), torch.randn(K, device=DEV)
    out = torch.zeros(rows, 6 * K, device=DEV, dtype=torch.bfloat16)
    ops.split3(x, out, mode, g if mode == 2 else None, b if mode == 2 else None, 1e-5)
    seg = out.view(rows, 6, K)
    assert torch.equal(seg[:, 0], seg[:, 3]) and torch.equal(seg[:, 2], seg[:, 4]) and torch.equal(seg[:, 2], seg[:, 5])
    val = seg[:, 2].float() + seg[:, 0].float() + seg[:, 1].float()          # h + m + l
    if mode == 0:
        pre = x
        h, m, l = _split3_ref(x)
        assert torch.equal(seg[:, 2], h) and torch.equal(seg[:, 0], m) and torch.equal(seg[:, 1], l)
    elif mode == 1:
        pre = torch.nn.functional.gelu(x.double()).float()
    else:
        pre = torch.nn.functional.layer_norm(x.double(), (K,), g.double(), b.double(), 1e-5).float()
    assert rel_err(val, pre) < 2e-6


@pytest.mark.parametrize("M,N,K", [(300, 256, 192), (1000, 768, 3072)])
def test_split_gemm_reaches_fp32_accuracy(ops, M, N, K):
    """One bf16 tcgen05 GEMM over the six split products == the fp32 product (float64 reference), where the plain bf16 GEMM
    is three decimal digits away."""
    from peekvit_b200._lib import PK_EPI_BIAS_F32
    a = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV) / math.sqrt(K)
    bias = torch.randn(N, device=DEV) * 0.1
    ref = a.double() @ w.double().t() + bias.double()
    a6 = ops.split3(a, torch.empty(M, 6 * K, device=DEV, dtype=torch.bfloat16))
    out = ops.gemm(a6, ops.split3_weight(w), bias, torch.empty(M, N, device=DEV), PK_EPI_BIAS_F32)
    assert rel_err(out.double(), ref) < TOL_EXACT
    plain = ops.gemm(a.to(torch.bfloat16), w.to(torch.bfloat16), bias, torch.empty(M, N, device=DEV), PK_EPI_BIAS_F32)
    assert rel_err(plain.double(), ref) > 50 * TOL_EXACT


@pytest.mark.parametrize("B,H,dh,n", [(3, 2, 64, 197), (2, 8, 32, 785), (5, 3, 64, 17), (2, 2, 32, 130), (2, 8, 48, 197)])
def test_attention_f32(ops, B, H, dh, n):
    D = H * dh
    qkv = torch.randn(B * n, 3 * D, device=DEV)
    out = ops.attention_f32(qkv, torch.zeros(B * n, D, device=DEV), B, H, dh, n)
    q, k, v = (t.reshape(B, n, H, dh).transpose(1, 2).double() for t in qkv.view(B, n, 3 * D).split(D, dim=-1))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1) @ v).transpose(1, 2).reshape(B * n, D)
    assert rel_err(out.double(), ref) < 5e-6          # fp32 accumulation over up to 785 keys against a float64 reference


def test_patchify_split3(ops):
    img = torch.randn(3, 3, 48, 48, device=DEV)
    P, Kp = 9, 3 * 16 * 16
    out = ops.patchify_split3(img, 16, torch.zeros(3 * P, 6 * Kp, device=DEV, dtype=torch.bfloat16))
    ref = torch.nn.functional.unfold(img, kernel_size=16, stride=16).transpose(1, 2).reshape(3 * P, Kp)
    h, m, l = _split3_ref(ref)
    seg = out.view(3 * P, 6, Kp)
    assert torch.equal(seg[:, 2], h) and torch.equal(seg[:, 0], m) and torch.equal(seg[:, 1], l)
    assert torch.equal(seg[:, 3], m) and torch.equal(seg[:, 4], h) and torch.equal(seg[:, 5], h)


def test_head_as_split_gemm_matches_fused_head(ops):
    """cls_features + split3 + GEMM (the large-batch head) == cls_head (fused CUDA-core kernel) == torch fp32."""
    from peekvit_b200._lib import PK_EPI_BIAS_F32
    B, seq, D, C, T = 300, 21, 256, 1000, 2
    x = torch.randn(B * seq, D, device=DEV)
    g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    w, hb = torch.randn(C, D, device=DEV) * 0.05, torch.randn(C, device=DEV)
    fused = ops.cls_head(x, B, seq, T, g, b, 1e-5, w, hb)
    feat = ops.cls_features(x, B, seq, T, g, b, 1e-5, torch.empty(B, D, device=DEV))
    f6 = ops.split3(feat, torch.empty(B, 6 * D, device=DEV, dtype=torch.bfloat16))
    got = ops.gemm(f6, ops.split3_weight(w), hb, torch.empty(B, C, device=DEV), PK_EPI_BIAS_F32)
    ref = torch.nn.functional.linear(torch.nn.functional.layer_norm(x.view(B, seq, D)[:, :T].double(), (D,), g.double(), b.double(), 1e-5).sum(1),
                                     w.double(), hb.double())
    assert rel_err(got.double(), ref) < 5e-6 and rel_err(fused.double(), ref) < 5e-6
    cu = (torch.arange(B + 1, device=DEV, dtype=torch.int32) * seq)
    feat2 = ops.cls_features(x, B, 0, T, g, b, 1e-5, torch.empty(B, D, device=DEV), cu_seqlens=cu)
    assert torch.equal(feat, feat2)


# ------------------------------------------------------------------ ragged tcgen05 attention (csrc/pk_attention_tc.cu, "tcr")
def _ragged_case(lens, H, seed, with_mult, with_extra, pad_rows=0):
    dh, D = 64, H * 64
    g = torch.Generator(device=DEV).manual_seed(seed)
    rows = sum(lens)
    qkv = torch.zeros(rows + pad_rows, 3 * D, device=DEV, dtype=torch.bfloat16)
    qkv[:rows] = (torch.randn(rows, 3 * D, device=DEV, generator=g) * 1.2).to(torch.bfloat16)
    cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), device=DEV, dtype=torch.int32)
    km = ekv = em = None
    if with_mult:
        km = torch.randint(1, 60, (rows + pad_rows,), device=DEV, generator=g).float()
        km[torch.rand(rows + pad_rows, device=DEV, generator=g) < 0.7] = 1.0          # mostly plain keys, like the ghost-row layout
    if with_extra:
        ekv = (torch.randn(2 * D, device=DEV, generator=g) * 0.5).to(torch.bfloat16)
        em = torch.randint(0, 150, (len(lens),), device=DEV, generator=g).float()
        em[::3] = 0.0
    return qkv, cu, km, ekv, em, rows


@pytest.mark.parametrize("with_mult,with_extra", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("lens", [[3, 70, 198, 1, 129, 64], [199, 199, 16, 17, 31, 32, 33, 127, 128, 1], [255, 200, 130, 5], [12] * 9,
                                  [80, 77, 91, 64, 85, 79, 102, 60, 66, 71, 93, 88]])
def test_attention_ragged_tcgen05_against_reference(ops, lens, with_mult, with_extra):
    """impl=3 forces the ragged tcgen05 kernel: samples of unequal length (one or two 128-query tiles, key counts on both sides
    of every 16-key group boundary), per-key multiplicities, the virtual bias key, rows of the buffer past the live ones; rows
    past the live ones stay untouched.  Compared with the fp32 reference and with the general mma.sync kernel (impl=1)."""
    B, H, dh = len(lens), 6, 64
    D = H * dh
    if max(lens) + (1 if with_extra else 0) > 256:
        pytest.skip("more than 256 keys")
    qkv, cu, km, ekv, em, rows = _ragged_case(lens, H, 7 * sum(lens) + with_mult + 2 * with_extra, with_mult, with_extra, pad_rows=40)
    out = torch.full((rows + 40, D), 3.0, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, H, dh, cu_seqlens=cu, max_seq_len=max(lens), key_mult=km, extra_kv=ekv, extra_mult=em, impl=3)
    assert ops.device_flag() == 0
    ref = ref_attention(qkv[:rows], B, H, dh, lens, km[:rows] if km is not None else None, ekv, em)
    assert rel_err(out[:rows], ref) < TOL_BF16
    assert bool((out[rows:] == 3.0).all())
    out1 = torch.zeros_like(out)
    ops.attention(qkv, out1, B, H, dh, cu_seqlens=cu, max_seq_len=max(lens), key_mult=km, extra_kv=ekv, extra_mult=em, impl=1)
    assert rel_err(out[:rows], out1[:rows]) < TOL_BF16


@pytest.mark.parametrize("N", [1, 5, 14, 17, 26, 33, 50, 64, 99, 128, 129, 197, 256])
def test_attention_ragged_tcgen05_uniform_lengths(ops, N):
    """The same kernel on uniform samples (what the pruned RankViT layers run: 5 / 14 / 26 / 50 tokens)."""
    B, H, dh = 7, 12, 64
    D = H * dh
    qkv = (torch.randn(B * N, 3 * D, device=DEV) * 1.5).to(torch.bfloat16)
    out = torch.full((B * N, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, H, dh, seq_len=N, impl=3)
    assert ops.device_flag() == 0
    assert rel_err(out, ref_attention(qkv, B, H, dh, [N] * B)) < TOL_BF16


def test_attention_ragged_tcgen05_full_size_row_sum_property(ops):
    """ResidualViT-S at budget 0.4 size (512 images x 6 heads, ~80 live rows each, multiplicities and virtual keys): with
    V = ones and a virtual value of ones every output element is exactly 1, whatever Q, K and the multiplicities are."""
    B, H, dh = 512, 6, 64
    D = H * dh
    g = torch.Generator().manual_seed(3)
    lens = torch.randint(20, 200, (B,), generator=g).tolist()
    qkv, cu, km, ekv, em, rows = _ragged_case(lens, H, 11, True, True, pad_rows=300)
    qkv[:, 2 * D:] = 1.0
    ekv[D:] = 1.0
    out = torch.zeros(rows + 300, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, H, dh, cu_seqlens=cu, max_seq_len=max(lens), key_mult=km, extra_kv=ekv, extra_mult=em, impl=3)
    assert ops.device_flag() == 0
    assert float((out[:rows].float() - 1.0).abs().max()) < 1e-2
    assert bool((out[rows:] == 0).all())


@pytest.mark.parametrize("min_rows", [1, 10 ** 9])
def test_attention_ragged_device_side_routing(ops, min_rows):
    """route_rows / route_min_rows: both ragged kernels are launched and the device-side live row count picks the one that
    runs (threshold 1: the tcgen05 kernel; threshold 1e9: the general kernel); the other leaves the output untouched."""
    lens = [150, 33, 198, 77, 120]
    B, H, dh = len(lens), 6, 64
    D = H * dh
    qkv, cu, km, ekv, em, rows = _ragged_case(lens, H, 5, True, True, pad_rows=16)
    out = torch.full((rows + 16, D), 3.0, device=DEV, dtype=torch.bfloat16)
    rows_dev = torch.tensor([rows], device=DEV, dtype=torch.int32)
    ops.attention(qkv, out, B, H, dh, cu_seqlens=cu, max_seq_len=199, key_mult=km, extra_kv=ekv, extra_mult=em,
                  route_rows=rows_dev, route_min_rows=min_rows)
    assert ops.device_flag() == 0
    ref = ref_attention(qkv[:rows], B, H, dh, lens, km[:rows], ekv, em)
    assert rel_err(out[:rows], ref) < TOL_BF16
    assert bool((out[rows:] == 3.0).all())
    forced = torch.zeros_like(out)
    ops.attention(qkv, forced, B, H, dh, cu_seqlens=cu, max_seq_len=199, key_mult=km, extra_kv=ekv, extra_mult=em,
                  impl=3 if min_rows == 1 else 1)
    assert torch.equal(out[:rows], forced[:rows])          # bit-identical to the kernel the threshold selects


# ------------------------------------------------------------------ quad-region ragged tcgen05 attention ("tcq", <= 128 keys per sample)
@pytest.mark.parametrize("with_mult,with_extra", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("lens", [[3, 70, 127, 1, 64, 65, 33, 16, 17], [80, 77, 91, 64, 85, 79, 102, 60, 66, 71, 93, 88] * 3,
                                  [12] * 9, [127, 127, 127, 127, 5], [40, 0, 23, 0, 0, 90, 1]])
def test_attention_quad_region_tcgen05_against_reference(ops, lens, with_mult, with_extra):
    """impl=4 forces the quad-region kernel: four samples in flight per SM, one (sample, head) per unit, key counts on both
    sides of every 16-key group boundary, per-key multiplicities, the virtual bias key, EMPTY samples (an A-ViT sample whose
    class token has halted), rows of the buffer past the live ones left untouched."""
    B, H, dh = len(lens), 6, 64
    D = H * dh
    if max(lens) + (1 if with_extra else 0) > 128:
        pytest.skip("more than 128 keys")
    qkv, cu, km, ekv, em, rows = _ragged_case(lens, H, 5 * sum(lens) + with_mult + 2 * with_extra, with_mult, with_extra, pad_rows=40)
    out = torch.full((rows + 40, D), 3.0, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, H, dh, cu_seqlens=cu, max_seq_len=max(lens), key_mult=km, extra_kv=ekv, extra_mult=em, impl=4)
    assert ops.device_flag() == 0
    ref = ref_attention(qkv[:rows], B, H, dh, lens, km[:rows] if km is not None else None, ekv, em)
    assert rel_err(out[:rows], ref) < TOL_BF16
    assert bool((out[rows:] == 3.0).all())


@pytest.mark.parametrize("lens", [[100, 33, 77, 5, 64], [190, 33, 127, 128, 0, 129, 64], [197, 150, 129, 199], [128, 127, 1, 140, 90, 16]])
@pytest.mark.parametrize("tcr_long", [False, True])
def test_attention_per_sample_split_between_ragged_kernels(ops, lens, tcr_long, monkeypatch):
    """PK_ATT_SPLIT=1 (opt-in, read per call): one ragged call, split per sample on the device: the quad-region tcgen05 kernel takes the samples of at most 128 keys
    (rows + the virtual key), the two-region tcgen05 kernel (tcr_long) or the general mma.sync kernel the longer ones; each
    sample's rows are bit-identical to the kernel that owns it launched alone, and the whole is the reference's."""
    monkeypatch.setenv("PK_ATT_SPLIT", "1")
    B, H, dh = len(lens), 6, 64
    D = H * dh
    qkv, cu, km, ekv, em, rows = _ragged_case(lens, H, 9, True, True, pad_rows=16)
    tot = torch.tensor([rows], device=DEV, dtype=torch.int32)
    out = torch.full((rows + 16, D), 3.0, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, H, dh, cu_seqlens=cu, max_seq_len=199, key_mult=km, extra_kv=ekv, extra_mult=em,
                  route_rows=tot, route_min_rows=0 if tcr_long else 10 ** 9)
    assert ops.device_flag() == 0
    ref = ref_attention(qkv[:rows], B, H, dh, lens, km[:rows], ekv, em)
    assert rel_err(out[:rows], ref) < TOL_BF16 and bool((out[rows:] == 3.0).all())
    alone = torch.zeros_like(out)
    ops.attention(qkv, alone, B, H, dh, cu_seqlens=cu, max_seq_len=199, key_mult=km, extra_kv=ekv, extra_mult=em,
                  impl=3 if tcr_long else 1)
    keys = [n + (1 if e > 0 else 0) for n, e in zip(lens, em.tolist())]      # extra_mult <= 0: no virtual key for that sample
    start = 0
    for n, k in zip(lens, keys):
        if k > 128:                          # a long sample: the other kernel's, exactly
            assert torch.equal(out[start:start + n], alone[start:start + n])
        start += n
    short = [n if k <= 128 else 0 for n, k in zip(lens, keys)]
    if any(short):                           # the short samples: exactly the quad-region kernel's rows
        idx = torch.cat([torch.arange(s0, s0 + n) for s0, n in zip(cu.tolist(), short)]).to(DEV)
        cu_s = torch.tensor([0] + list(torch.tensor(short).cumsum(0)), device=DEV, dtype=torch.int32)
        qkv_s = torch.zeros(idx.numel() + 128, 3 * D, device=DEV, dtype=torch.bfloat16)
        qkv_s[:idx.numel()] = qkv[idx]
        km_s = torch.ones(idx.numel() + 128, device=DEV, dtype=torch.float32)
        km_s[:idx.numel()] = km[idx]
        out_s = torch.zeros(idx.numel() + 128, D, device=DEV, dtype=torch.bfloat16)
        ops.attention(qkv_s, out_s, B, H, dh, cu_seqlens=cu_s, max_seq_len=127, key_mult=km_s, extra_kv=ekv, extra_mult=em, impl=4)
        assert torch.equal(out[idx], out_s[:idx.numel()])
