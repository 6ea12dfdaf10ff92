"""The "bf16x2" arithmetic mode (model.pk_precision = "bf16x2"): every linear operand is carried as hi + lo (two bf16 terms,
16 significant bits; one tcgen05 GEMM over K' = 3K with the activation row stored once as [lo | hi]), the split is produced by
the GEMM / attention epilogues themselves, and the tcgen05 attention runs on IEEE-half operands.  It is the mode that meets
all three north-star numbers at once: logits within 1e-2 (measured ~2e-4), top-1 agreement >= 99.9 %, CUDA-graph replayed.
Kernel-level checks first (against float64 products), then the model against the oracle / the fp32-accurate mode."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
VITB = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


@pytest.mark.parametrize("M,N,K", [(1000, 768, 768), (300, 1536, 384), (5, 256, 128), (2500, 768, 3072)])
def test_split2_gemm_against_float64(M, N, K):
    """A = [lo | hi] read with a_wrap_k against W' = [Wh | Wl | Wh]: 16-bit operands, fp32 accumulation."""
    from peekvit_b200 import ops
    from peekvit_b200._lib import PK_EPI_BIAS_F32
    a = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    bias = torch.randn(N, device=DEV)
    a2 = ops.split(a, torch.empty(M, 2 * K, dtype=torch.bfloat16, device=DEV), 2)
    hi = a.to(torch.bfloat16)
    assert torch.equal(a2[:, K:], hi) and torch.equal(a2[:, :K], (a - hi.float()).to(torch.bfloat16))
    out = ops.gemm(a2, ops.split_weight(w, 2), bias, torch.empty(M, N, device=DEV), PK_EPI_BIAS_F32, a_wrap_k=K, cta_pair=2)
    ref = a.double() @ w.double().T + bias.double()
    plain = ops.gemm(hi, w.to(torch.bfloat16), bias, torch.empty(M, N, device=DEV), PK_EPI_BIAS_F32)
    err, err_plain = _rel(out, ref), _rel(plain, ref)
    print(f"split2 GEMM {M}x{N}x{K}: {err:.2e} of max|out| (plain bf16 operands {err_plain:.2e})")
    assert err < 2e-5 and err_plain > 20 * err
    assert ops.device_flag() == 0


def test_gemm_half_and_split_outputs():
    """The two epilogue variants of the mode: IEEE-half output (in-projection -> attention operands) and the GELU output
    written already split [lo | hi] (fc1 -> fc2 operand), full tiles and a ragged row count."""
    from peekvit_b200 import ops
    from peekvit_b200._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_GELU_BF16, PK_OUT_BF16X2, PK_OUT_F16
    for M in (777, 40):
        K, N = 384, 1536
        a = torch.randn(M, K, device=DEV)
        w = torch.randn(N, K, device=DEV) / K ** 0.5
        bias = torch.randn(N, device=DEV) * 0.1
        a2 = ops.split(a, torch.empty(M, 2 * K, dtype=torch.bfloat16, device=DEV), 2)
        w3 = ops.split_weight(w, 2)
        ref = a.double() @ w.double().T + bias.double()
        half = ops.gemm(a2, w3, bias, torch.zeros(M, N, dtype=torch.float16, device=DEV), PK_EPI_BIAS_BF16, a_wrap_k=K, cta_pair=2,
                        out_format=PK_OUT_F16)
        assert _rel(half, ref) < 1e-3                      # half rounding of the result: 2^-11 relative per element
        assert torch.equal(half, (ref.float()).to(torch.float16)) or (half.float() - ref.float()).abs().max() < 2e-3 * ref.abs().max()
        m_dev = torch.tensor([M - 3], device=DEV, dtype=torch.int32)
        out = torch.full((M, 2 * N), 7.0, dtype=torch.bfloat16, device=DEV)
        ops.gemm(a2, w3, bias, out, PK_EPI_BIAS_GELU_BF16, a_wrap_k=K, cta_pair=2, out_format=PK_OUT_BF16X2, m_dev=m_dev)
        g = torch.nn.functional.gelu(ref)
        got = out[:M - 3, N:].double() + out[:M - 3, :N].double()
        assert ((got - g[:M - 3]).abs().max() / g.abs().max()).item() < 3e-5
        hi_p, lo_p = out[:M - 3, N:].float(), out[:M - 3, :N].float()
        assert bool((lo_p.abs() <= hi_p.abs() * 2.0 ** -8 + 1e-30).all())          # lo is the rounding residual of hi: at most half an ulp
        assert bool((out[M - 3:] == 7.0).all())            # rows past the device-side count stay untouched
    assert ops.device_flag() == 0


@pytest.mark.parametrize("n", [197, 198, 50, 99])
def test_half_attention_with_split_output(n):
    from peekvit_b200 import ops
    B, H, dh = 5, 6, 64
    D = H * dh
    qkv = (torch.randn(B * n, 3 * D, device=DEV) * 0.7)
    q16 = qkv.to(torch.float16)
    out = ops.attention(q16, torch.empty(B * n, 2 * D, dtype=torch.bfloat16, device=DEV), B, H, dh, seq_len=n, half_split=True)
    x = q16.double().view(B, n, 3, H, dh)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    ref = (torch.softmax(q @ k.transpose(-1, -2) / dh ** 0.5, -1) @ v).transpose(1, 2).reshape(B * n, D)
    got = out[:, D:].double() + out[:, :D].double()
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    bf = ops.attention(q16.float().to(torch.bfloat16), torch.empty(B * n, D, dtype=torch.bfloat16, device=DEV), B, H, dh, seq_len=n)
    err_bf = ((bf.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"half attention n={n}: {err:.2e} (bf16 kernel on the same values {err_bf:.2e})")
    assert err < 1.5e-3 and err < err_bf
    assert ops.device_flag() == 0


def test_vit_b16_bf16x2_meets_all_three_north_star_numbers():
    """Logits within 1e-2 of the fp32 oracle (asserted at 1e-3) on 64 images; top-1 agreement >= 99.9 % on 4096 images against
    the fp32-accurate mode (itself within 1e-5 of the oracle and 100 % top-1, asserted in test_models_gpu.py); and the mode is
    at least 2.2x faster than the fp32-accurate one (>= 8.5k img/s on a B200, the torch-eager bf16 figure, is recorded by
    bench.py's precision_modes block)."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200 import ops
    from peekvit_b200.models import VisionTransformer
    sd = ow.make_state_dict("vit", VITB, seed=4321)
    model = VisionTransformer(**VITB)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    images = ow.synthetic_images(64, 224, seed=77)
    ref = torch.cat([po.forward("vit", sd, VITB, images[s:s + 32])[0] for s in range(0, 64, 32)])
    model.pk_precision = "bf16x2"
    logits = model(images.to(DEV)).cpu()
    err = _rel(logits, ref)
    print(f"bf16x2 ViT-B/16: {err:.2e} of max|logit| against the fp32 oracle")
    assert err < 1e-3
    assert torch.equal(logits.argmax(1), ref.argmax(1))
    g = torch.Generator(device=DEV).manual_seed(4096)
    big = torch.randn(4096, 3, 224, 224, device=DEV, generator=g)
    t = {}
    outs = {}
    for mode in ("fp32", "bf16x2", "bf16"):
        model.pk_precision = mode
        model(big)                                   # same batch once untimed: graphs captured, workspaces allocated
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(2):                           # best of two: a parity suite must not fail on one slow pass
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            outs[mode] = model(big)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        t[mode] = best
    assert ops.device_flag() == 0
    exact = outs["fp32"]
    agree = {m: (outs[m].argmax(1) == exact.argmax(1)).float().mean().item() for m in ("bf16x2", "bf16")}
    errs = {m: _rel(outs[m], exact) for m in ("bf16x2", "bf16")}
    print(f"4096 images: top-1 agreement with the fp32 mode bf16x2 {agree['bf16x2']:.5f} (bf16 {agree['bf16']:.5f}); "
          f"logits {errs['bf16x2']:.2e} (bf16 {errs['bf16']:.2e}); images/s fp32 {4096 / t['fp32'] * 1e3:.0f}, "
          f"bf16x2 {4096 / t['bf16x2'] * 1e3:.0f}, bf16 {4096 / t['bf16'] * 1e3:.0f}")
    assert errs["bf16x2"] < 1e-3
    assert agree["bf16x2"] >= 0.999
    assert t["bf16x2"] < t["fp32"] / 1.8            # measured 2.9 - 3.0x (profiles/r02); the throughput itself is bench.py's business


@pytest.mark.parametrize("name", ["vit_d64_h2", "vit_d128_regs", "rankvit_b05", "residual_learnable_cal04", "avit", "moevit", "moevit_attn",
                                  "eeresidual_learnable_cal04", "residual_skip_attention", "residual_skip_mlp_add_input",
                                  "residual_skip_mlp_fixed", "residual_two_cls_cal05"])
def test_bf16x2_mode_matches_reference_fixture(name):
    """Every family runs in the mode (ragged / small-head shapes take the fp32 attention core): logits within 1e-3 of the
    reference fixture -- 10x inside the bf16 tolerance."""
    import os
    import numpy as np
    from golden_cases import CASES, build_case
    from peekvit_b200 import ops, runner
    from peekvit_b200.models import build_model
    names = {"vit": "vit", "rankvit": "RankVisionTransformer", "residualvit": "residualvit", "adavit": "adavit", "moevit": "vitmoe",
             "eeresidualvit": "eeResidualvit"}
    case = CASES[name]
    sd, images = build_case(case)
    model = build_model(names[case["family"]], case["cfg"])
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    if case.get("budget") is not None:
        model.set_budget(case["budget"])
    model.pk_precision = "bf16x2"
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    out = model(images.to(DEV))
    logits = (out[-1] if isinstance(out, list) else out).cpu().numpy()
    assert ops.device_flag() == 0
    err = np.abs(logits - ref["logits"]).max() / np.abs(ref["logits"]).max()
    print(f"bf16x2 {name}: {err:.2e}")
    assert err < 1e-3
    again = model(images.to(DEV))                     # second call replays the CUDA graph
    again = (again[-1] if isinstance(again, list) else again).cpu().numpy()
    assert np.array_equal(again, logits)
