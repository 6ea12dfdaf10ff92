"""Model-level GPU parity: the drop-in modules (C-ABI CUDA path) against (a) the fixtures the
imported reference produced (tests/golden/*.npz) and (b) the CPU oracle on the same seeded
weights and images.  North-star tolerances: bf16 logits within 1e-2 of max|ref|; kept-token index
sets bit-exact given identical fp32 scores (ties -> lowest index)."""
import os

import numpy as np
import pytest
import torch

from golden_cases import CASES, build_case, oracle_noise

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_LOGITS = 1e-2
GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = {"vit": "vit", "rankvit": "RankVisionTransformer", "residualvit": "residualvit", "adavit": "adavit", "moevit": "vitmoe",
         "eeresidualvit": "eeResidualvit"}
BUILT = ("vit", "rankvit", "residualvit", "adavit", "moevit", "eeresidualvit")


def _model(case):
    from peekvit_b200.models import build_model
    sd, images = build_case(case)
    model = build_model(NAMES[case["family"]], case["cfg"])
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    if case.get("budget") is not None:
        model.set_budget(case["budget"])
    return model, sd, images


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if CASES[n].get("noise") is not None])
def test_noise_block_matches_oracle_and_fixture(name):
    """NoiseBlock spliced in by add_noise (utils/utils.py:162-191).  The draw comes from torch's generators like in the
    reference: re-seeding and drawing the same tensor first gives the oracle the exact noise the CUDA path used.  Token
    dropping draws its permutation on the host generator, so it also reproduces the reference fixture itself."""
    from oracle import peekvit_oracle as po
    from peekvit_b200 import ops
    from peekvit_b200.models import add_noise, build_model, NoiseBlock
    case = CASES[name]
    model, sd, images = _model(case)
    clean = model(images.to(DEV)).cpu()
    nb = add_noise(model, **case["noise"])
    assert isinstance(nb, NoiseBlock) and model.encoder.layers[case["noise"]["layer"]] is nb
    torch.manual_seed(case["noise_seed"])
    logits = model(images.to(DEV)).cpu()
    assert ops.device_flag() == 0
    residual = case["family"] == "residualvit"
    with torch.no_grad():
        if residual:
            ref, raux = po.residualvit_forward(sd, case["cfg"], images, case["budget"], noise=oracle_noise(case, device=DEV))
        else:
            ref, _ = po.vit_forward(sd, case["cfg"], images, noise=oracle_noise(case, device=DEV))
    scale = ref.abs().max()
    assert ((logits - ref).abs().max() / scale).item() < TOL_LOGITS
    assert ((logits - clean).abs().max() / scale).item() > 5 * TOL_LOGITS          # the noise really acted
    if case["noise"]["noise_type"] == "token_drop":
        gold = np.load(os.path.join(GOLD, name + ".npz"))["logits"]
        assert np.abs(logits.numpy() - gold).max() / np.abs(gold).max() < TOL_LOGITS
    if residual:
        # block.mask of the layers behind the NoiseBlock (encoder.layers index = block index + 1 there)
        blocks = [b for b in model.encoder.layers if not isinstance(b, NoiseBlock)]
        for j, blk in enumerate(blocks):
            m, g = blk.mask.cpu(), raux["masks"][j]
            assert m.shape == g.shape and (m - g).abs().max().item() < 5e-3
    nb.set_value(0.0)                                   # 0 dB / probability 0 switch the block off (blocks.py:125-127,:145)
    off = model(images.to(DEV)).cpu()
    if residual:
        # with the block spliced in the model runs on the dense masked row layout, without it on compacted rows: the same
        # function, two summation orders
        assert ((off - clean).abs().max() / scale).item() < TOL_LOGITS
    else:
        assert torch.equal(off, clean)
    # standalone call on a (B, N, D) tensor
    x = torch.randn(2, 7, 64, device=DEV)
    g = NoiseBlock("gaussian", snr=5.0)
    torch.manual_seed(3)
    y = g(x)
    torch.manual_seed(3)
    n = torch.randn(2, 7, 64, device=DEV)
    exp = x + n * torch.sqrt((x ** 2).mean(-1, keepdim=True) / 10 ** 0.5)
    assert torch.allclose(y, exp, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if CASES[n]["family"] in BUILT and CASES[n].get("noise") is None])
def test_model_matches_reference_fixture(name):
    from peekvit_b200 import ops, runner
    case = CASES[name]
    model, sd, images = _model(case)
    aux = {}
    ref = np.load(os.path.join(GOLD, name + ".npz"))
    if case["family"] == "eeresidualvit":
        # list output (eeresidualvit.py:355-357): L early exits, then the final logits
        outs = model(images.to(DEV))
        L = case["cfg"]["num_layers"]
        assert isinstance(outs, list) and len(outs) == L + 1
        for i in range(L):
            r = ref[f"exit_{i}"]
            assert tuple(outs[i].shape) == r.shape
            assert np.abs(outs[i].cpu().numpy() - r).max() / np.abs(r).max() < TOL_LOGITS
        logits = outs[-1].cpu().numpy()
        # the host-resident entry returns the same list (host tensors)
        host = model.forward_host(images.pin_memory())
        assert len(host) == L + 1 and all(h.device.type == "cpu" for h in host)
        assert all(torch.allclose(h, o.cpu(), atol=1e-6) for h, o in zip(host, outs))
    else:
        logits = runner.run(model, images.to(DEV), aux).cpu().numpy()
    assert ops.device_flag() == 0
    scale = np.abs(ref["logits"]).max()
    if case["cfg"].get("gate_type") == "gumbel":
        # hard 0 / 1 gates (round(sigmoid), blocks.py:55-57): a bf16-level difference in a gate logit next to 0 flips a whole
        # token, so -- like the RankViT selections -- the logits are compared given identical decisions: the oracle replayed
        # with the masks this path published (the decisions themselves are compared with the fixture below, and are identical
        # in the fp32 mode: test_fp32_mode_matches_reference_fixture)
        from oracle import peekvit_oracle as po
        with torch.no_grad():
            replay, _ = po.residualvit_forward(sd, case["cfg"], images, case.get("budget"),
                                               forced_masks={i: torch.cat(m).cpu() for i, m in aux["masks"].items()})
        assert np.abs(logits - replay.numpy()).max() / scale < TOL_LOGITS
    else:
        # the 'attention' / 'mlp' skip modes return mlp(...) WITHOUT the residual around it (residualvit.py:155-157,186-187):
        # the stream is then a bare bf16-operand GEMM output, re-normalised by the next LayerNorm, instead of a small update
        # on an fp32 stream -- operand rounding shows up at a few 1e-2 over four such layers.  The same cases are held to
        # 1e-3 in the bf16x2 mode (tests/test_bf16x2_gpu.py) and to 1e-5 in the fp32 mode (below).
        no_resid = any(m in ("attention", "mlp") for m in (case["cfg"].get("residual_layers") or []))
        assert np.abs(logits - ref["logits"]).max() / scale < (4 * TOL_LOGITS if no_resid else TOL_LOGITS)
        assert (logits.argmax(1) == ref["logits"].argmax(1)).mean() >= 0.75   # 3-4 images: at most one near-tie flip
    if case["family"] == "rankvit":
        assert aux["seq_lens"] == list(ref["seq_lens"])
        for i, kept in aux["kept"].items():
            # exact given OUR scores (the contract) ...
            exp = torch.argsort(aux["scores"][i], dim=-1, descending=True, stable=True)[:, :kept.shape[1]]
            assert torch.equal(kept.long(), exp)
            # ... and the same token sets as the fp32 reference up to bf16-induced near-ties at the cut
            g = ref[f"kept_{i}"]
            overlap = np.mean([len(set(a) & set(b)) / len(a) for a, b in zip(kept.cpu().numpy().tolist(), g.tolist())])
            assert overlap >= 0.9
    if case["family"] in ("residualvit", "eeresidualvit"):
        # published side-state: block.mask (B, N_img, 1) soft gate values (utils/utils.py:100-122).  Values are
        # continuous in the activations (bf16 band); a keep/drop flag can only differ at a near-tie with the
        # threshold, where the soft value itself is ~0.
        agree = []
        for i, blk in enumerate(model.encoder.layers):
            if f"mask_{i}" not in ref:                 # plain layers (skip None / 'none') publish no mask
                continue
            g = torch.from_numpy(ref[f"mask_{i}"])
            m = blk.mask.cpu()
            assert m.shape == g.shape
            if case["cfg"].get("gate_type") != "gumbel":           # hard gates: a near-tie flip is a difference of 1
                assert (m - g).abs().max().item() < 5e-3
            agree.append(((m > 0) == (g > 0)).float().mean().item())
        assert np.mean(agree) >= 0.9
    if case["family"] == "adavit":
        assert model.encoder.rho_token.shape == ref["rho_token"].shape
        model.pk_early_exit = False           # exact per-token ACT bookkeeping needs every sample run to the end
        logits2 = runner.run(model, images.to(DEV)).cpu().numpy()
        assert np.abs(logits2 - ref["logits"]).max() / scale < TOL_LOGITS
        cnt = model.encoder.counter_token.cpu().numpy()
        assert (cnt == ref["counter_token"]).mean() >= 0.97        # a halting near-tie may move a token by one layer
    if case["family"] == "moevit":
        for i, blk in enumerate(model.encoder.layers):
            if blk.mlp.gating_probs is not None:
                gp = blk.mlp.gating_probs
                assert gp.shape[-1] == blk.mlp.num_experts and torch.all(gp.sum(-1) == 1)
                assert (gp.argmax(-1).cpu().numpy() == ref[f"mlp_gating_{i}"]).mean() >= 0.98
            gp = blk.self_attention.gating_probs
            if blk.self_attention.num_experts > 1:
                assert gp.shape[-1] == blk.self_attention.num_experts and torch.all(gp.sum(-1) == 1)
                assert (gp.argmax(-1).cpu().numpy() == ref[f"attn_gating_{i}"]).mean() >= 0.98


def test_vit_tiny_config_a_against_oracle():
    """BASELINE config A: vit_tiny p8 D256 H8(dh32) F768 L4 @224 -> 785 tokens, 10 classes."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200.models import VisionTransformer
    cfg = dict(image_size=224, patch_size=8, num_layers=4, num_heads=8, hidden_dim=256, mlp_dim=768, num_classes=10)
    sd = ow.make_state_dict("vit", cfg, seed=4321)
    images = ow.synthetic_images(16, 224, seed=1234)
    ref, _ = po.forward("vit", sd, cfg, images)
    model = VisionTransformer(**cfg)
    model.load_state_dict(sd)
    logits = model.to(DEV).eval()(images.to(DEV)).cpu()
    assert ((logits - ref).abs().max() / ref.abs().max()).item() < TOL_LOGITS


def test_vit_b16_against_oracle_and_batch_invariance():
    """BASELINE config B shape on 24 images: parity with the oracle, and logits independent of
    micro-batch split (samples are independent: the property sharding across GPUs relies on)."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200.models import VisionTransformer
    cfg = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
    sd = ow.make_state_dict("vit", cfg, seed=4321)
    images = ow.synthetic_images(24, 224, seed=1234)
    ref, _ = po.forward("vit", sd, cfg, images)
    model = VisionTransformer(**cfg)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    x = images.to(DEV)
    logits = model(x)
    assert ((logits.cpu() - ref).abs().max() / ref.abs().max()).item() < TOL_LOGITS
    assert (logits.cpu().argmax(1) == ref.argmax(1)).float().mean().item() >= 0.95
    model.pk_micro_batch = 5
    assert torch.equal(model(x), logits)
    assert torch.equal(model(x[7:9]), logits[7:9])


def test_rankvit_b16_budget_sweep_against_oracle():
    """BASELINE config D: RankViT on the ViT-B shape, rank layers [3,6,9], budget 0.5 then 0.25."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200 import runner
    from peekvit_b200.models import RankVisionTransformer
    cfg = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000,
               rankvit_layers=[3, 6, 9])
    sd = ow.make_state_dict("rankvit", cfg, seed=4321)
    images = ow.synthetic_images(8, 224, seed=1234)
    model = RankVisionTransformer(**cfg)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    for budget, lens in ((0.5, [197] * 3 + [99] * 3 + [50] * 3 + [26] * 3), (0.25, [197] * 3 + [50] * 3 + [14] * 3 + [5] * 3),
                         (1.0, [197] * 12)):
        model.set_budget(budget)
        aux = {}
        logits = runner.run(model, images.to(DEV), aux).cpu()
        assert aux["seq_lens"] == lens
        kept = {i: k.cpu() for i, k in aux.get("kept", {}).items()}
        for i, k in kept.items():       # bit-exact stable top-k of our own fp32 scores
            exp = torch.argsort(aux["scores"][i], dim=-1, descending=True, stable=True)[:, :k.shape[1]]
            assert torch.equal(k.long(), exp.cpu())
        # Logits given identical selections: random-init token norms are nearly equal, so bf16-level score
        # noise swaps tokens at the cut w.r.t. the fp32 oracle (top-k is discontinuous); the oracle is
        # therefore replayed with the selections the CUDA path made.
        ref, oaux = po.rankvit_forward(sd, cfg, images, budget, forced_kept=kept)
        err = ((logits - ref).abs().max() / ref.abs().max()).item()
        free, _ = po.rankvit_forward(sd, cfg, images, budget)
        err_free = ((logits - free).abs().max() / free.abs().max()).item()
        if kept:
            first = min(kept)
            sc = oaux["scores"][first]
            assert ((aux["scores"][first].cpu() - sc).abs().max() / sc.abs().max()).item() < TOL_LOGITS
        print(f"rankvit budget {budget}: rel err {err:.3e} (given selections), {err_free:.3e} (oracle's own selections)")
        assert err < TOL_LOGITS


def test_module_contract_on_device():
    from peekvit_b200.models import VisionTransformer
    model = VisionTransformer(image_size=32, patch_size=8, num_layers=2, num_heads=2, hidden_dim=64, mlp_dim=128, num_classes=10)
    with torch.no_grad():       # random-init head / class token are zero (vit.py:165,186-188): make logits non-vacuous
        model.head.weight.normal_(std=0.1)
        model.class_tokens.normal_(std=0.5)
    model = model.to(DEV)
    x = torch.randn(2, 3, 32, 32, device=DEV)
    with pytest.raises(RuntimeError):          # training mode: inference path only
        model(x)
    model.eval()
    assert model(x).shape == (2, 10)
    with pytest.raises(AssertionError):        # wrong image size (torch._assert, vit.py:206-207)
        model(torch.randn(2, 3, 40, 40, device=DEV))
    with pytest.raises(RuntimeError):          # CPU input to a CUDA model
        model(torch.randn(2, 3, 32, 32))
    # prepacked weights follow in-place parameter updates and layer deletion
    y0 = model(x)
    assert y0.abs().max() > 1e-3
    with torch.no_grad():
        model.head.bias.add_(1.0)
    assert torch.allclose(model(x), y0 + 1.0, atol=1e-6)
    del model.encoder.layers[1]                # topology.py:177-178 / vit.py:312-313
    y1 = model(x)
    assert y1.shape == (2, 10) and not torch.allclose(y1, y0 + 1.0, atol=1e-4)
    # host-resident batch: same logits through the double-buffered H2D path
    yh = model.forward_host(x.cpu().pin_memory())
    assert yh.device.type == "cpu" and torch.allclose(yh, y1.cpu(), atol=1e-6)


def test_residualvit_really_compacts_and_scales_with_budget():
    """ViT-S shape (BASELINE config C), calibrated gates: the packed row count per layer follows the budget,
    logits stay within the bf16 band of the oracle, and masks are published per layer."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200 import runner
    from peekvit_b200.models import ResidualVisionTransformer
    cfg = dict(image_size=224, patch_size=16, num_layers=12, num_heads=6, hidden_dim=384, mlp_dim=1536, num_classes=1000,
               gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
               residual_layers=["attention+mlp"] * 12)
    sd = ow.make_state_dict("residualvit", cfg, seed=4321)
    sd = ow.calibrate_residual_gates(sd, cfg, 0.4, images=ow.synthetic_images(2, 224, seed=99))
    images = ow.synthetic_images(6, 224, seed=1234)
    model = ResidualVisionTransformer(**cfg)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    model.set_budget(0.4)
    aux = {}
    logits = runner.run(model, images.to(DEV), aux).cpu()
    ref, oaux = po.forward("residualvit", sd, cfg, images, 0.4)
    err = ((logits - ref).abs().max() / ref.abs().max()).item()
    rows = [int(r[0][0]) for _, r in sorted(aux["rows"].items())]
    keep = [float((m > 0).float().mean()) for _, m in sorted(oaux["masks"].items())]
    print(f"residualvit ViT-S budget 0.4: rel err {err:.3e}; packed rows/layer {rows} (dense {6 * 198}); oracle keep {keep}")
    assert err < TOL_LOGITS
    assert max(rows) <= 6 * 199 and np.mean(rows) < 0.75 * 6 * 198          # survivors only
    for i, blk in enumerate(model.encoder.layers):
        assert blk.mask.shape == (6, 196, 1)
        assert (blk.mask.cpu() - oaux["masks"][i]).abs().max().item() < 5e-3


def test_short_sequence_models_run_on_the_quad_region_attention(monkeypatch):
    """160-px images, patch 16: at most 103 rows per sample (+ the virtual key), so the static bound sends every ragged attention
    call of ResidualViT and A-ViT to the quad-region tcgen05 kernel alone (head_dim 64).  Logits within the bf16 band of the
    oracle, masks / halting counters as in the 224-px tests; PK_ATT_TCQ is not touched, i.e. this is the default dispatch."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200 import ops, runner
    from peekvit_b200.models import AdaptiveVisionTransformer, ResidualVisionTransformer
    base = dict(image_size=160, patch_size=16, num_layers=6, num_heads=4, hidden_dim=256, mlp_dim=512, num_classes=100)
    images = ow.synthetic_images(9, 160, seed=77)
    cfg = dict(base, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
               residual_layers=["attention+mlp"] * 6)
    sd = ow.calibrate_residual_gates(ow.make_state_dict("residualvit", cfg, seed=11), cfg, 0.5, images=ow.synthetic_images(2, 160, seed=5))
    model = ResidualVisionTransformer(**cfg)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    model.set_budget(0.5)
    aux = {}
    logits = runner.run(model, images.to(DEV), aux).cpu()
    ref, oaux = po.forward("residualvit", sd, cfg, images, 0.5)
    err = ((logits - ref).abs().max() / ref.abs().max()).item()
    rows = [int(r[0][0]) for _, r in sorted(aux["rows"].items())]
    print(f"residualvit 160 px budget 0.5: rel err {err:.3e}; packed rows/layer {rows} (dense {9 * 102})")
    assert ops.device_flag() == 0 and err < TOL_LOGITS
    assert 0 < min(rows) and np.mean(rows) < 0.9 * 9 * 102
    for i, blk in enumerate(model.encoder.layers):
        assert (blk.mask.cpu() - oaux["masks"][i]).abs().max().item() < 5e-3
    # A-ViT: halting is discontinuous (a counter that moves by one layer re-weighs that token's output), so the bf16 band of
    # this family is the config-E test's 5e-2; the same model with the quad-region kernel switched off (general mma.sync
    # kernel; a fresh module, i.e. a fresh captured graph) must land in the same band with the same counters
    cfg = dict(base, eps=0.01, gate_scale=1.0, gate_center=1.5)
    sd = ow.make_state_dict("adavit", cfg, seed=12)
    ref, oaux = po.forward("adavit", sd, cfg, images)
    got = {}
    for tcq in ("1", "0"):
        monkeypatch.setenv("PK_ATT_TCQ", tcq)
        model = AdaptiveVisionTransformer(**cfg)
        model.load_state_dict(sd)
        model = model.to(DEV).eval()
        logits = runner.run(model, images.to(DEV)).cpu()
        err = ((logits - ref).abs().max() / ref.abs().max()).item()
        cnt = model.encoder.counter_token.cpu()
        agree = (cnt == oaux["counter_token"]).float().mean().item()
        print(f"adavit 160 px (PK_ATT_TCQ={tcq}): rel err {err:.3e}; mean layers per token {float(cnt.mean()):.2f}; counter agreement {agree:.4f}")
        assert ops.device_flag() == 0 and err < 5e-2 and agree > 0.98
        assert float(cnt.mean()) < 5.9          # tokens really halt
        got[tcq] = (logits, cnt)
    assert (got["1"][1] == got["0"][1]).float().mean().item() > 0.99
    assert ((got["1"][0] - got["0"][0]).abs().max() / ref.abs().max()).item() < 5e-2


def test_vit_b16_top1_agreement_outside_the_tolerance_band():
    """North star: logits within 1e-2 relative and top-1 agreement >= 99.9 %.  With random-init weights the fp32 top-2
    margins are tiny (median 0.15 against max|logit| 2.7), so a few arg-max flips inside the bf16 error band are
    unavoidable for any bf16-operand path; the checkable statement is: every sample whose fp32 margin exceeds the
    tolerance band (2 x 1e-2 x max|logit|) has the same top-1, and overall agreement stays above 97 %."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200.models import VisionTransformer
    cfg = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
    sd = ow.make_state_dict("vit", cfg, seed=4321)
    images = ow.synthetic_images(256, 224, seed=77)
    ref = torch.cat([po.forward("vit", sd, cfg, images[s:s + 32])[0] for s in range(0, 256, 32)])
    model = VisionTransformer(**cfg)
    model.load_state_dict(sd)
    logits = model.to(DEV).eval()(images.to(DEV)).cpu()
    scale = ref.abs().max()
    assert ((logits - ref).abs().max() / scale).item() < TOL_LOGITS
    top2 = ref.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    same = logits.argmax(1) == ref.argmax(1)
    clear = margin > 2 * TOL_LOGITS * scale
    assert bool(same[clear].all()), "top-1 differs on a sample whose fp32 margin is outside the bf16 tolerance band"
    assert same.float().mean().item() >= 0.97


def test_uint8_input_path_matches_float_path():
    """uint8 HWC images with ToTensor + Normalize fused into the im2col give bit-identical patches and
    logits to the float path fed with the tensor the torchvision transforms produce (data/imagenette.py:69-73)."""
    from peekvit_b200 import ops
    from peekvit_b200.models import VisionTransformer
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (6, 64, 64, 3), generator=g, dtype=torch.uint8)
    mean, std = torch.tensor(ops.IMAGENET_MEAN), torch.tensor(ops.IMAGENET_STD)
    flt = ((u8.permute(0, 3, 1, 2).float().div(255) - mean[None, :, None, None]) / std[None, :, None, None]).contiguous()
    p_u8 = ops.patchify_u8(u8.to(DEV), 16)
    p_f = ops.patchify(flt.to(DEV), 16)
    assert torch.equal(p_u8, p_f)
    model = VisionTransformer(image_size=64, patch_size=16, num_layers=2, num_heads=2, hidden_dim=128, mlp_dim=256, num_classes=10)
    with torch.no_grad():
        model.head.weight.normal_(std=0.1)
        model.class_tokens.normal_(std=0.5)
    model = model.to(DEV).eval()
    y_f = model(flt.to(DEV))
    assert y_f.abs().max() > 1e-3
    assert torch.equal(model(u8.to(DEV)), y_f)
    assert torch.equal(model.forward_host(u8.pin_memory()), y_f.cpu())
    with pytest.raises(AssertionError):
        model(torch.zeros(2, 3, 64, 64, dtype=torch.uint8, device=DEV))       # uint8 must be HWC


# ------------------------------------------------------------------ fp32-accurate mode
TOL_FP32_MODE = 1e-5        # north star: logits within 1e-5 relative in fp32 mode


def test_forward_host_prefetch_of_the_next_batch_changes_nothing_but_the_schedule():
    """forward_host(x, next_host=y): the first chunk of y is staged while x still computes.  Three different batches in a row
    (prefetched, prefetched, last), then a call whose tensor is NOT the one that was announced, then a batch of another size:
    every result equals the plain call's, bit for bit."""
    from oracle import weights as ow
    from peekvit_b200.models import VisionTransformer
    cfg = dict(image_size=64, patch_size=8, num_layers=2, num_heads=2, hidden_dim=128, mlp_dim=256, num_classes=10)
    model = VisionTransformer(**cfg)
    model.load_state_dict(ow.make_state_dict("vit", cfg, seed=5))
    model = model.to(DEV).eval()
    model.pk_micro_batch = 160                      # 300 images: chunks [40, 120, 140]
    batches = [ow.synthetic_images(300, 64, seed=s).pin_memory() for s in (1, 2, 3, 4)]
    plain = [model.forward_host(b).clone() for b in batches]
    got = [model.forward_host(batches[0], next_host=batches[1]).clone(),
           model.forward_host(batches[1], next_host=batches[2]).clone(),
           model.forward_host(batches[2], next_host=batches[0]).clone(),       # announces batch 0 ...
           model.forward_host(batches[3]).clone()]                              # ... but batch 3 arrives
    for a, b in zip(plain, got):
        assert torch.equal(a, b)
    small = ow.synthetic_images(50, 64, seed=9).pin_memory()
    ref_small = model.forward_host(small).clone()
    model.forward_host(batches[0], next_host=small)                             # another shape: not staged
    assert torch.equal(model.forward_host(small, next_host=small), ref_small)
    assert torch.equal(model.forward_host(small), ref_small)                    # staged by the call before (same tensor)
    u8 = (torch.rand(200, 64, 64, 3) * 255).to(torch.uint8).pin_memory()
    ref_u8 = model.forward_host(u8).clone()
    model.forward_host(u8, next_host=u8)
    assert torch.equal(model.forward_host(u8), ref_u8)


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if CASES[n].get("noise") is None])
def test_fp32_mode_matches_reference_fixture(name):
    """model.pk_precision = 'fp32': split-operand tcgen05 GEMMs + fp32 attention reproduce the reference's fp32 logits to
    1e-5 for every family, and with them the reference's discrete decisions: kept-token indices (RankViT), keep / drop
    flags and soft mask values (ResidualViT, EE-ResidualViT), halting counters (A-ViT), expert routing (MoE)."""
    from peekvit_b200 import ops, runner
    case = CASES[name]
    fam = case["family"]
    model, sd, images = _model(case)
    model.pk_precision = "fp32"
    ref = np.load(os.path.join(GOLD, name + ".npz"))
    scale = np.abs(ref["logits"]).max()
    aux = {}
    if fam == "eeresidualvit":
        outs = model(images.to(DEV))
        for i in range(case["cfg"]["num_layers"]):
            r = ref[f"exit_{i}"]
            assert np.abs(outs[i].cpu().numpy() - r).max() / np.abs(r).max() < TOL_FP32_MODE
        logits = outs[-1].cpu().numpy()
    else:
        if fam == "adavit":
            model.pk_early_exit = False           # per-token ACT bookkeeping of every sample to the last layer
        logits = runner.run(model, images.to(DEV), aux).cpu().numpy()
    assert ops.device_flag() == 0
    assert np.abs(logits - ref["logits"]).max() / scale < TOL_FP32_MODE
    if fam == "rankvit":
        assert aux["seq_lens"] == list(ref["seq_lens"])
        for i, kept in aux["kept"].items():
            assert np.array_equal(kept.cpu().numpy(), ref[f"kept_{i}"])          # bit-exact index sets, in order
    if fam in ("residualvit", "eeresidualvit"):
        for i, blk in enumerate(model.encoder.layers):
            if f"mask_{i}" not in ref:
                continue
            g = ref[f"mask_{i}"]
            m = blk.mask.cpu().numpy()
            assert np.array_equal(m > 0, g > 0)                                  # identical keep / drop decisions
            assert np.abs(m - g).max() < 2e-6
    if fam == "adavit":
        assert np.array_equal(model.encoder.counter_token.cpu().numpy(), ref["counter_token"])
        assert np.allclose(model.encoder.rho_token.cpu().numpy(), ref["rho_token"], atol=1e-5)
    if fam == "moevit":
        for i, blk in enumerate(model.encoder.layers):
            for moe, key in ((blk.mlp, "mlp_gating"), (blk.self_attention, "attn_gating")):
                if moe.num_experts > 1:
                    assert np.array_equal(moe.gating_probs.argmax(-1).cpu().numpy(), ref[f"{key}_{i}"])
    if fam in ("vit", "rankvit"):
        model.pk_precision = "bf16"
        again = runner.run(model, images.to(DEV)).cpu().numpy()
        assert 1e-4 < np.abs(again - ref["logits"]).max() / scale < TOL_LOGITS


def test_fp32_mode_config_a_and_vit_b16_against_oracle():
    """BASELINE config A (vit_tiny p8, 785 tokens, dh 32, the reference's CPU-runnable fp32 case) and the ViT-B/16 shape."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200.models import VisionTransformer
    for cfg, n_img in ((dict(image_size=224, patch_size=8, num_layers=4, num_heads=8, hidden_dim=256, mlp_dim=768, num_classes=10), 8),
                       (dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000), 6)):
        sd = ow.make_state_dict("vit", cfg, seed=4321)
        images = ow.synthetic_images(n_img, 224, seed=1234)
        ref, _ = po.forward("vit", sd, cfg, images)
        model = VisionTransformer(**cfg)
        model.load_state_dict(sd)
        model = model.to(DEV).eval()
        model.pk_precision = "fp32"
        logits = model(images.to(DEV)).cpu()
        assert ((logits - ref).abs().max() / ref.abs().max()).item() < TOL_FP32_MODE
        assert torch.equal(logits.argmax(1), ref.argmax(1))


def test_precision_flag_is_validated():
    case = CASES["vit_d64_h2"]
    model, sd, images = _model(case)
    model.pk_precision = "fp16"
    with pytest.raises(ValueError):
        model(images.to(DEV))


# ------------------------------------------------------------------ the reference's shipped model configs
_REF_CONFIGS = {
    # configs/model/<name>.yaml of the reference: (class alias, kwargs); image 224 unless the config is an Imagenette one
    "deit_t_16_224": ("vit", dict(patch_size=16, num_layers=12, hidden_dim=192, mlp_dim=768, num_heads=3)),
    "deit_s_16_224": ("vit", dict(patch_size=16, num_layers=12, hidden_dim=384, mlp_dim=1536, num_heads=6)),
    "vit_small": ("vit", dict(patch_size=16, num_layers=8, hidden_dim=384, mlp_dim=1536, num_heads=8)),          # head_dim 48
    "vit_tiny": ("vit", dict(patch_size=8, num_layers=4, hidden_dim=256, mlp_dim=768, num_heads=8)),              # head_dim 32
    "rankvit": ("rankvit", dict(patch_size=8, num_layers=4, hidden_dim=256, mlp_dim=768, num_heads=4, rankvit_layers=[1, 2, 3])),
    "rankdeit_t_16_224": ("rankvit", dict(patch_size=16, num_layers=12, hidden_dim=192, mlp_dim=768, num_heads=3,
                                          rankvit_layers=[3, 6, 9])),
}


@pytest.mark.parametrize("name", sorted(_REF_CONFIGS))
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_reference_model_configs_run_and_match_the_oracle(name, precision):
    """Every dense / rank model shape the reference ships a config for (configs/model/*.yaml): head_dim 32 / 48 / 64,
    hidden 192 - 384, patch 8 / 16, at 224 px; both arithmetic modes against the fp32 oracle."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200.models import build_model
    fam, kw = _REF_CONFIGS[name]
    cfg = dict(kw, image_size=224, num_classes=10)
    sd = ow.make_state_dict(fam, cfg, seed=77)
    images = ow.synthetic_images(3, 224, seed=78)
    budget = 0.5 if fam == "rankvit" else None
    ref, _ = po.forward(fam, sd, cfg, images, budget)
    model = build_model(NAMES[fam], cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    if budget is not None:
        model.set_budget(budget)
    model.pk_precision = precision
    logits = model(images.to(DEV)).cpu()
    err = ((logits - ref).abs().max() / ref.abs().max()).item()
    # RankViT in bf16 mode may swap tokens at the cut (discontinuous selection): its band is the fixture tests' one
    assert err < (TOL_FP32_MODE if precision == "fp32" else (5e-2 if fam == "rankvit" else TOL_LOGITS)), err


# ------------------------------------------------------------------ evaluation loop (validate/test.py:97-156)
def test_evaluate_loop_accuracy_and_analytic_cost():
    """peekvit_b200.evaluate: per-budget accuracy with the count kept on the device, cost per image from the token counts
    of the forward (against the oracle's predictions and the oracle's FLOP formula on its own kept-token counts)."""
    from oracle import peekvit_oracle as po
    from peekvit_b200.evaluate import evaluate
    case = CASES["rankvit_b05"]
    model, sd, images = _model(case)
    model.pk_precision = "fp32"                     # exact selections -> exact comparisons
    ref, oaux = po.forward("rankvit", sd, case["cfg"], images, 0.5)
    labels = ref.argmax(1)
    labels[0] = (labels[0] + 1) % case["cfg"]["num_classes"]            # one deliberately wrong label
    batches = [(images[:3], labels[:3]), (images[3:], labels[3:])]
    res = evaluate(model, batches, budgets=[1.0, 0.5])
    assert set(res) == {1.0, 0.5}
    assert res[0.5]["accuracy"] == pytest.approx((len(images) - 1) / len(images))
    assert res[0.5]["tokens_per_layer"] == [float(n) for n in oaux["seq_lens"]]
    assert res[0.5]["gmacs_per_image"] == pytest.approx(po.flops_per_image(case["cfg"], oaux["seq_lens"]) / 2e9, rel=1e-6)
    assert res[1.0]["gmacs_per_image"] == pytest.approx(po.flops_per_image(case["cfg"]) / 2e9, rel=1e-6)
    assert res[0.5]["gmacs_per_image"] < res[1.0]["gmacs_per_image"] and res[0.5]["images_per_second"] > 0
    # ResidualViT: live-row counts stay on the device until the end of the pass; fewer tokens at the smaller budget
    rcase = CASES["residual_learnable_cal04"]
    rmodel, rsd, rimg = _model(rcase)
    rres = evaluate(rmodel, [(rimg, torch.zeros(len(rimg), dtype=torch.long))], budgets=[0.9, 0.2])
    for b in (0.9, 0.2):
        assert len(rres[b]["tokens_per_layer"]) == rcase["cfg"]["num_layers"] and rres[b]["gmacs_per_image"] > 0
    # reference accounting of the last pass (budget 0.2): class + budget token + the tokens the published masks keep; the
    # rows actually computed are at most that + one ghost row (re-admitted dropped tokens travel as one row)
    for i, blk in enumerate(rmodel.encoder.layers):
        kept = (blk.mask[..., 0] > 0).sum(1).float()
        assert rres[0.2]["tokens_per_layer"][i] == pytest.approx(float((2 + kept).mean()))
        assert rres[0.2]["computed_rows_per_layer"][i] <= rres[0.2]["tokens_per_layer"][i] + 1
    # noise sweep (validate/test.py:108-112): results keyed by noise value
    from peekvit_b200.models import add_noise
    vcase = CASES["vit_d64_h2"]
    vmodel, vsd, vimg = _model(vcase)
    nb = add_noise(vmodel, layer=1, noise_type="gaussian", snr=10.0)
    vres = evaluate(vmodel, [(vimg, torch.zeros(len(vimg), dtype=torch.long))], noise_module=nb, noise_vals=[0.0, 5.0])
    assert set(vres[None]) == {0.0, 5.0} and "accuracy" in vres[None][5.0]


def test_evaluate_loop_other_families():
    """A-ViT (live-row lists per micro-batch), EE-ResidualViT (list output -> final head) and MoE (dense accounting)."""
    from peekvit_b200.evaluate import evaluate
    for name in ("avit", "eeresidual_learnable_cal04", "moevit"):
        case = CASES[name]
        model, sd, images = _model(case)
        model.pk_micro_batch = 3                    # two micro-batches: the per-layer counts are summed over both
        labels = torch.zeros(len(images), dtype=torch.long)
        res = evaluate(model, [(images, labels)], budgets=[case.get("budget")])[case.get("budget")]
        assert 0.0 <= res["accuracy"] <= 1.0 and res["gmacs_per_image"] > 0
        L = case["cfg"]["num_layers"]
        if name == "moevit":
            assert res["tokens_per_layer"] is None
        else:
            assert len(res["tokens_per_layer"]) == L
            full = 65 + (1 if name.startswith("eeresidual") else 0)
            assert all(0 <= t <= full for t in res["tokens_per_layer"]), res["tokens_per_layer"]
        if name == "avit":
            # halting removes tokens: later layers see fewer rows than the first
            assert res["tokens_per_layer"][-1] < res["tokens_per_layer"][0] == 65


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if CASES[n].get("noise") is None])
def test_single_image_and_empty_batches(name):
    """Edge cases of the batch dimension for every family: one image (every GEMM falls back to the single-CTA kernel) gives
    the logits it gets inside the batch; an empty batch returns an empty result instead of launching anything."""
    from peekvit_b200 import ops
    case = CASES[name]
    if case["family"] == "residualvit" and case["cfg"].get("add_budget_token") is True:
        pytest.skip("a fixed-float budget token thresholds on the batch mean: single-image results differ by construction")
    model, sd, images = _model(case)
    x = images.to(DEV)
    full = model(x)
    one = model(x[2:3])
    none = model(x[:0])
    if isinstance(full, list):
        full, one, none = full[-1], one[-1], none[-1]
    assert ops.device_flag() == 0
    assert tuple(one.shape) == (1, case["cfg"]["num_classes"]) and tuple(none.shape) == (0, case["cfg"]["num_classes"])
    assert ((one - full[2:3]).abs().max() / full.abs().max()).item() < TOL_LOGITS
