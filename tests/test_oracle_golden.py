"""The oracle restatement replays the fixtures generated from the imported reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from golden_cases import CASES, build_case, oracle_noise
from oracle import peekvit_oracle as po

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_fixture(name):
    case = CASES[name]
    ref = np.load(os.path.join(GOLD, name + ".npz"))
    sd, images = build_case(case)
    if case.get("noise") is not None:
        # NoiseBlock cases: the reference drew from torch's CPU generator under noise_seed; the same draw is regenerated here
        with torch.no_grad():
            if case["family"] == "residualvit":
                logits, aux = po.residualvit_forward(sd, case["cfg"], images, case["budget"], noise=oracle_noise(case))
            else:
                logits, aux = po.vit_forward(sd, case["cfg"], images, noise=oracle_noise(case))
    else:
        logits, aux = po.forward(case["family"], sd, case["cfg"], images, case.get("budget"))
    if case["family"] == "eeresidualvit":
        # list output: one early exit per layer, then the final logits (eeresidualvit.py:355-357)
        assert len(logits) == case["cfg"]["num_layers"] + 1
        for i, e in enumerate(logits[:-1]):
            r = ref[f"exit_{i}"]
            assert e.shape == r.shape and np.abs(e.numpy() - r).max() / np.abs(r).max() < 1e-5
        logits = logits[-1]
    scale = np.abs(ref["logits"]).max()
    assert scale > 0.1, "vacuous fixture (zero logits)"
    # fp32 tolerance of the north star: 1e-5 relative
    assert np.abs(logits.numpy() - ref["logits"]).max() / scale < 1e-5
    fam = case["family"]
    if fam == "rankvit":
        for i, idx in aux["kept"].items():
            assert np.array_equal(idx.numpy().astype(np.int32), ref[f"kept_{i}"])   # bit-exact index sets, in order
        assert list(ref["seq_lens"]) == aux["seq_lens"]
    if fam in ("residualvit", "eeresidualvit"):
        for i, m in aux["masks"].items():
            assert m.shape == ref[f"mask_{i}"].shape                                   # (B, N_img, 1)
            assert np.allclose(m.numpy(), ref[f"mask_{i}"], atol=2e-6)
            assert np.array_equal(m.numpy() > 0, ref[f"mask_{i}"] > 0)
    if fam == "adavit":
        assert np.allclose(aux["rho_token"].numpy(), ref["rho_token"], atol=1e-5)
        assert np.array_equal(aux["counter_token"].numpy(), ref["counter_token"])
    if fam == "moevit":
        for i, gp in aux["mlp_gating"].items():
            assert np.array_equal(gp.argmax(-1).numpy().astype(np.int32), ref[f"mlp_gating_{i}"])
        for i, gp in aux["attn_gating"].items():
            assert np.array_equal(gp.argmax(-1).numpy().astype(np.int32), ref[f"attn_gating_{i}"])


def test_stable_topk_tie_rule():
    s = torch.tensor([[1.0, 3.0, 3.0, 0.0, 3.0, 1.0], [2.0] * 6])
    idx = po.stable_topk_desc(s, 3)
    assert idx.tolist() == [[1, 2, 4], [0, 1, 2]]
    z = torch.tensor([[0.0, -0.0, 0.0, 1e-45, -0.0]])
    assert po.stable_topk_desc(z, 2).tolist() == [[3, 0]]


def test_rank_keep_counts_compound():
    # 64 patches, rank layers [1,3], budget .5 -> 65, 33, 33, 17
    n = 64
    lens = []
    for i in range(4):
        if i in (1, 3):
            n = po.rank_keep_count(n, 0.5)
        lens.append(n + 1)
    assert lens == [65, 33, 33, 17]
    assert po.rank_keep_count(196, 0.5) == 98 and po.rank_keep_count(98, 0.5) == 49 and po.rank_keep_count(49, 0.5) == 25


def test_flops_formula_matches_survey():
    cfg = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
    assert abs(po.flops_per_image(cfg) / 1e9 - 35.128) < 0.01
    assert abs(po.flops_per_image(cfg, [197] * 3 + [99] * 3 + [50] * 3 + [26] * 3) / 1e9 - 16.508) < 0.01
