"""GPU parity at the shapes BASELINE.json names (configs C, D, E), against the CPU oracle on the same seeded weights and
images.  Gates are calibrated (tests/golden/calibration.json, written by tests/golden/make_calibration.py with the oracle) so
that the budgets are realised instead of the keep-everything / keep-nothing behaviour of random gates.

Tolerances (north star): bf16 logits within 1e-2 of max|logit|; discrete decisions (keep / drop flags, halting counters,
expert routing) may differ only at near-ties, so they are bounded by an agreement fraction; the fp32-accurate mode must
reproduce them exactly."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_LOGITS = 1e-2
GOLD = os.path.join(os.path.dirname(__file__), "golden")
VITS = dict(image_size=224, patch_size=16, num_layers=12, num_heads=6, hidden_dim=384, mlp_dim=1536, num_classes=1000)
VITB = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
CFG_C = dict(VITS, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
             add_input=False, residual_layers=["attention+mlp"] * 12)          # configs/model/residualdeit_s_16_224.yaml


def _calibration():
    with open(os.path.join(GOLD, "calibration.json")) as f:
        return json.load(f)


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("budget", [0.2, 0.4, 0.8, 1.0])
def test_config_c_residualvit_s_every_budget(budget):
    """BASELINE config C: ResidualViT at the ViT-S shape with the learnable budget token, budgets 0.2 / 0.4 / 0.8 / 1.0:
    logits, published masks and the number of rows actually computed per layer, in the three arithmetic modes.
    A keep / drop flag is a threshold decision on a continuous score: it may differ from the fp32 oracle's only where the
    score is within the arithmetic's noise of the threshold: the soft mask VALUES must match the oracle's within 1e-2 (bf16),
    1e-4 (bf16x2), 5e-6 (fp32 mode), flags may differ only where both values are below that band, and the fp32-accurate mode
    reproduces every flag."""
    from make_calibration import config_c_state_dict
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200 import ops, runner
    from peekvit_b200.models import ResidualVisionTransformer
    cal = _calibration()
    sd = config_c_state_dict(cal["seed"], cal["gate_std"], cal["bt_gate_scale"])
    for i, b in enumerate(cal["config_C"][str(budget)]):
        sd[f"encoder.layers.{i}.residual_gate.projection.bias"] = torch.tensor([b])
    images = ow.synthetic_images(6, 224, seed=1234)
    ref, oaux = po.forward("residualvit", sd, CFG_C, images, budget)
    model = ResidualVisionTransformer(**CFG_C)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    model.set_budget(budget)
    keep_ref = [float((m > 0).float().mean()) for _, m in sorted(oaux["masks"].items())]
    mean_keep = float(np.mean(keep_ref))
    assert abs(mean_keep - min(budget, 0.97)) < 0.25                     # the calibration realises the budget (6 test images)

    def run(mode, band):
        """-> logits error, flag agreement, max |mask - ref|, rows computed per layer, tokens kept per layer.  The soft mask
        relu(score - threshold) is continuous in the activations, so the statement that holds for every arithmetic is: mask
        VALUES within ``band`` of the oracle's; a keep / drop FLAG can then differ only where both values are below ``band``
        (a token kept with a mask of ~0 contributes ~0: the logits stay continuous across the flip)."""
        model.pk_precision = mode
        aux = {}
        logits = runner.run(model, images.to(DEV), aux).cpu()
        assert ops.device_flag() == 0
        agree, dmax, kept = [], 0.0, []
        for i, blk in enumerate(model.encoder.layers):
            m, g = blk.mask.cpu(), oaux["masks"][i]
            assert m.shape == (6, 196, 1)
            dmax = max(dmax, (m - g).abs().max().item())
            flip = (m > 0) != (g > 0)
            assert bool((torch.maximum(m, g)[flip] < band).all())
            agree.append(1.0 - flip.float().mean().item())
            kept.append(int((m > 0).sum()))
        rows = [int(r[0][0]) for _, r in sorted(aux["rows"].items())]
        return _rel(logits, ref), float(np.mean(agree)), dmax, rows, kept

    err, agree, dmax, rows, kept = run("bf16", 1e-2)
    print(f"config C budget {budget}: oracle keep/layer {[round(k, 2) for k in keep_ref]} (mean {mean_keep:.3f}); rows/layer {rows} of {6 * 198}")
    print(f"  bf16: rel err {err:.3e}, flag agreement {agree:.4f}, max |mask - ref| {dmax:.2e}")
    assert err < TOL_LOGITS and dmax < 1e-2          # measured 5.5e-3 - 7.3e-3: the gate's wide projection (std 4) amplifies the bf16 band
    # survivors only: per layer, the rows computed are the kept tokens + class token, budget token and one ghost row per sample
    for r, k in zip(rows, kept):
        assert r <= k + 3 * 6
    assert np.mean(rows) < (0.9 if budget < 1.0 else 1.01) * 6 * 199
    err, agree, dmax, _, _ = run("bf16x2", 1e-4)
    print(f"  bf16x2: rel err {err:.3e}, flag agreement {agree:.5f}, max |mask - ref| {dmax:.2e}")
    assert err < 1e-3 and dmax < 1e-4
    err, agree, dmax, _, _ = run("fp32", 5e-6)
    assert err < 1e-5 and agree == 1.0 and dmax < 5e-6


def test_config_e_avit_s_against_oracle():
    """BASELINE config E: A-ViT at the ViT-S shape, avit_s_16_224.yaml kwargs (eps 0.01, gate_scale 10) with the calibrated
    halting centre.  The halting score is sigmoid(10 * x[..., 0] - centre) (adavit.py:74): the gate multiplies every
    perturbation of the activations by 10 before it weighs the block outputs, so bf16 operands land at ~2e-2 of max|logit| on
    this model (measured 1.0e-2 - 1.9e-2; bounded at 5e-2 here, halting counters >= 90 % identical and at most two layers apart); the bf16x2 mode
    is held to the north star's 1e-2 with a 10x margin and the fp32-accurate mode reproduces counters and rho exactly."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200 import ops, runner
    from peekvit_b200.models import AdaptiveVisionTransformer
    cal = _calibration()
    cfg = dict(VITS, eps=0.01, gate_scale=cal["config_E_avit"]["gate_scale"], gate_center=cal["config_E_avit"]["gate_center"])
    sd = ow.make_state_dict("adavit", cfg, seed=cal["seed"])
    images = ow.synthetic_images(6, 224, seed=1234)
    ref, oaux = po.forward("adavit", sd, cfg, images)
    cref = oaux["counter_token"]
    mean_layers = float(cref.mean())
    assert 3.0 < mean_layers < 11.0                                      # halting really happens, progressively
    model = AdaptiveVisionTransformer(**cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    for mode, tol, min_agree in (("bf16", 5e-2, 0.90), ("bf16x2", 1e-3, 0.995), ("fp32", 1e-5, 1.0)):
        model.pk_precision = mode
        model.pk_early_exit = True
        logits = runner.run(model, images.to(DEV)).cpu()
        assert ops.device_flag() == 0
        model.pk_early_exit = False
        logits2 = runner.run(model, images.to(DEV)).cpu()
        cnt = model.encoder.counter_token.cpu()
        agree = (cnt == cref).float().mean().item()
        print(f"config E A-ViT-S {mode}: rel err {_rel(logits, ref):.3e} (sample early exit) / {_rel(logits2, ref):.3e}; "
              f"mean layers per token {mean_layers:.2f}; counter agreement {agree:.4f}")
        assert _rel(logits, ref) < tol and _rel(logits2, ref) < tol
        assert agree >= min_agree and (cnt - cref).abs().max().item() <= {"bf16": 2, "bf16x2": 1, "fp32": 0}[mode]
    assert torch.allclose(model.encoder.rho_token.cpu(), oaux["rho_token"], atol=1e-5)


def test_config_e_moevit_s_four_experts_against_oracle():
    """BASELINE config E: VisionTransformerMoE at the ViT-S shape with mlp_moes=[4]*12 (attn_moes=None).  Arg-max routing is
    discontinuous (a near-tied router score moves a token to another expert's MLP), so with bf16 operands the logits are
    compared GIVEN the CUDA path's own routing (oracle replayed with it: within 1e-2) and free-running within 5e-2 with
    >= 97 % identical routing; bf16x2 free-running within 1e-2 and >= 99.9 %; the fp32-accurate mode routes every token like
    the oracle."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200 import ops, runner
    from peekvit_b200.models import VisionTransformerMoE
    cfg = dict(VITS, mlp_moes=[4] * 12)
    sd = ow.make_state_dict("moevit", cfg, seed=4321)
    images = ow.synthetic_images(6, 224, seed=1234)
    ref, oaux = po.forward("moevit", sd, cfg, images)
    model = VisionTransformerMoE(**cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    for mode, tol_free, min_agree in (("bf16", 5e-2, 0.97), ("bf16x2", 1e-2, 0.999), ("fp32", 1e-5, 1.0)):
        model.pk_precision = mode
        logits = runner.run(model, images.to(DEV)).cpu()
        assert ops.device_flag() == 0
        routed, agree = {}, []
        for i, blk in enumerate(model.encoder.layers):
            gp = blk.mlp.gating_probs
            assert gp.shape == (6, 197, 4) and torch.all(gp.sum(-1) == 1)
            routed[i] = gp.argmax(-1).cpu()
            agree.append((routed[i] == oaux["mlp_gating"][i].argmax(-1)).float().mean().item())
        with torch.no_grad():
            given, _ = po.moevit_forward(sd, cfg, images, forced_mlp_expert=routed)
        print(f"config E MoE-ViT-S {mode}: rel err {_rel(logits, ref):.3e} free-running, {_rel(logits, given):.3e} given its routing; "
              f"routing agreement per layer min {min(agree):.4f}")
        assert _rel(logits, ref) < tol_free and min(agree) >= min_agree
        assert _rel(logits, given) < (TOL_LOGITS if mode == "bf16" else tol_free)


@pytest.mark.parametrize("budget", [0.5, 0.4, 0.25])
def test_config_d_rankvit_b_free_running(budget):
    """BASELINE config D, free-running: the CUDA path makes its own selections and is compared with the oracle making ITS own.
    Top-k is discontinuous: random-init token norms are nearly equal, so where two norms at the cut differ by less than the
    arithmetic's noise the two paths keep different tokens, and a swapped token changes the logits by far more than the noise
    that swapped it.  Measured free-running error with bf16 operands: 4e-2 / 8e-2 / 1.6e-1 of max|logit| at budgets 0.5 / 0.4 /
    0.25 with 94 - 99.9 % of the kept sets shared; asserted: < 0.25 and >= 90 %.  (Given identical selections the error is
    < 1e-2: test_models_gpu.py::test_rankvit_b16_budget_sweep_against_oracle.)  The bf16x2 mode must do at least 4x better
    on the free-running error and the fp32-accurate mode keeps exactly the oracle's tokens with logits within 1e-5."""
    from oracle import peekvit_oracle as po, weights as ow
    from peekvit_b200 import runner
    from peekvit_b200.models import RankVisionTransformer
    cfg = dict(VITB, rankvit_layers=[3, 6, 9])
    sd = ow.make_state_dict("rankvit", cfg, seed=4321)
    images = ow.synthetic_images(8, 224, seed=1234)
    model = RankVisionTransformer(**cfg)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    model.set_budget(budget)
    free, oaux = po.rankvit_forward(sd, cfg, images, budget)
    res = {}
    for mode in ("bf16", "bf16x2", "fp32"):
        model.pk_precision = mode
        aux = {}
        logits = runner.run(model, images.to(DEV), aux).cpu()
        overlaps = []
        for i, kept in aux["kept"].items():
            a, b = kept.cpu().numpy().tolist(), oaux["kept"][i].numpy().tolist()
            overlaps.append(np.mean([len(set(x) & set(y)) / max(len(x), 1) for x, y in zip(a, b)]))
        res[mode] = (_rel(logits, free), min(overlaps), aux)
        print(f"config D budget {budget} {mode}: free-running rel err {res[mode][0]:.3e}; kept-set overlap per rank layer {[round(o, 4) for o in overlaps]}")
    assert res["bf16"][0] < 0.25 and res["bf16"][1] >= 0.90
    assert res["bf16x2"][0] < max(res["bf16"][0] / 4, 1e-3) and res["bf16x2"][1] >= 0.99
    assert res["fp32"][0] < 1e-5
    for i, kept in res["fp32"][2]["kept"].items():
        assert torch.equal(kept.cpu().long(), oaux["kept"][i])


def test_rankvit_budget_edge_cases():
    """rankvit.py:69-71: ceil(n*b) tokens are kept; a budget of 0 keeps the class token alone, a budget above 1 every token
    (same logits as budget 1: attention and the class readout do not depend on the token order)."""
    from golden_cases import CASES, build_case
    from oracle import peekvit_oracle as po
    from peekvit_b200 import runner
    from peekvit_b200.models import build_model
    case = CASES["rankvit_b05"]
    sd, images = build_case(case)
    model = build_model("RankVisionTransformer", case["cfg"])
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    model.set_budget(1.0)
    full = model(images.to(DEV)).cpu()
    model.set_budget(1.5)
    assert _rel(model(images.to(DEV)).cpu(), full) < TOL_LOGITS
    model.set_budget(0.0)
    aux = {}
    zero = runner.run(model, images.to(DEV), aux).cpu()
    ref, oaux = po.forward("rankvit", sd, case["cfg"], images, 0.0)
    assert aux["seq_lens"] == list(oaux["seq_lens"]) and aux["seq_lens"][-1] == 1
    assert _rel(zero, ref) < TOL_LOGITS


def test_topk_select_orders_nan_like_torch():
    """NaN scores sort above +inf (torch's descending argsort), ties to the lowest index: every slot of `kept` is written."""
    from peekvit_b200 import ops
    s = torch.tensor([[1.0, float("nan"), 3.0, float("nan"), float("inf"), -0.0, 0.0, -2.0]], device=DEV)
    kept = ops.topk_select(s, 8)
    assert kept.cpu().tolist() == [[1, 3, 4, 2, 0, 5, 6, 7]]
    assert torch.equal(kept.cpu().long(), torch.argsort(s.cpu(), dim=-1, descending=True, stable=True))
