"""The reference's OWN helper functions (utils/utils.py) run against the drop-in models: the unmodified reference tree
(``/root/reference`` in the authoring container, the vendored ``baseline/_ref/peekvit`` on the GPU box) is imported as package
``peekvit`` and only its ``peekvit.models`` sub-package is replaced by the B200 drop-ins (``install_as_peekvit``) -- what a
maintainer switching to this framework would do.  CPU part: checkpoint round trip, NoiseBlock splice, module introspection.
GPU part: the side-state the helpers read after a forward (block.mask, gating_probs)."""
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_BASE = dict(image_size=64, patch_size=8, num_layers=4, num_heads=2, hidden_dim=128, mlp_dim=256, num_classes=10)


def _reference_parent():
    """Directory that holds a ``peekvit`` package = the unmodified reference tree, or None."""
    vend = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(vend, "peekvit", "utils")):
        return vend
    if os.path.isdir("/root/reference/utils"):
        import tempfile
        tmp = tempfile.mkdtemp(prefix="peekvit_ref_")
        os.symlink("/root/reference", os.path.join(tmp, "peekvit"))
        return tmp
    return None


@pytest.fixture(scope="module")
def ref_utils():
    parent = _reference_parent()
    if parent is None:
        pytest.skip("no reference tree (baseline/_ref/peekvit or /root/reference)")
    saved = {k: v for k, v in sys.modules.items() if k == "peekvit" or k.startswith("peekvit.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, parent)
    try:
        import peekvit                                   # the reference tree (a namespace package)
        import peekvit_b200
        peekvit_b200.install_as_peekvit()                # peekvit.models.* -> the drop-ins
        utils = importlib.import_module("peekvit.utils.utils")
        assert "peekvit_b200" not in (getattr(utils, "__file__", "") or "")
        yield utils
    finally:
        sys.path.remove(parent)
        for k in [k for k in sys.modules if k == "peekvit" or k.startswith("peekvit.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_reference_checkpoint_helpers_and_introspection(ref_utils, tmp_path):
    """save_state / load_state (utils/utils.py:198-256), add_noise (:162-191), get_moes (:55-73), get_learned_thresholds (:125-135)."""
    from peekvit_b200.models import NoiseBlock, ResidualVisionTransformer, VisionTransformerMoE, build_model
    args = dict(_BASE, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
                residual_layers=["attention+mlp", None, "attention+mlp", "attention+mlp"])
    m = build_model("ResidualVisionTransformer", args)
    assert isinstance(m, ResidualVisionTransformer)
    exp_dir, ckpt_dir = ref_utils.make_experiment_directory(str(tmp_path / "exp"))
    ref_utils.save_state(ckpt_dir, m, args, None, None, epoch=7)
    ckpt = ref_utils.get_checkpoint_path(exp_dir, verbose=False)                         # 'last' epoch of the directory
    assert os.path.basename(ckpt) == "epoch_007.pth"
    orig_load = torch.load
    torch.load = lambda p, *a, **k: orig_load(p, *a, **{**k, "weights_only": False})     # the reference calls torch.load(path) bare
    try:
        m2, _, epoch, margs, nargs = ref_utils.load_state(ckpt, verbose=False)            # model=None: rebuilt through build_model
    finally:
        torch.load = orig_load
    assert epoch == 7 and margs["residual_layers"] == args["residual_layers"] and nargs is None
    assert isinstance(m2, ResidualVisionTransformer)
    sd, sd2 = m.state_dict(), m2.state_dict()
    assert list(sd) == list(sd2) and all(torch.equal(v, sd2[k]) for k, v in sd.items())
    thr = ref_utils.get_learned_thresholds(m)
    assert sorted(thr) == ["encoder.layers.0", "encoder.layers.2", "encoder.layers.3"] and all(t == 0.5 for t in thr.values())
    nb = ref_utils.add_noise(m, layer=2, noise_type="token_drop", prob=0.5)
    assert isinstance(nb, NoiseBlock) and m.encoder.layers[2] is nb
    moe = VisionTransformerMoE(**dict(_BASE, num_layers=3, mlp_moes=[1, 4, 2], attn_moes=[2, 1, 1]))
    assert sorted(ref_utils.get_moes(moe)) == ["encoder.layers.0.self_attention", "encoder.layers.1.mlp", "encoder.layers.2.mlp"]


@pytest.mark.gpu
def test_reference_side_state_helpers_after_a_forward(ref_utils):
    """get_forward_masks (utils/utils.py:100-122, plain and incremental) and get_last_forward_gates (:76-94) read what the
    drop-in forward published; values are checked against the oracle."""
    from golden_cases import CASES, build_case
    from oracle import peekvit_oracle as po
    from peekvit_b200.models import build_model
    case = CASES["residual_learnable_cal04"]
    sd, images = build_case(case)
    m = build_model("residualvit", case["cfg"])
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda:0").eval()
    m.set_budget(case["budget"])
    m.pk_precision = "fp32"
    m(images.to("cuda:0"))
    masks = ref_utils.get_forward_masks(m)
    _, aux = po.forward("residualvit", sd, case["cfg"], images, case["budget"])
    assert sorted(masks) == [f"encoder.layers.{i}" for i in range(4)]
    for i in range(4):
        got = masks[f"encoder.layers.{i}"].cpu()
        assert got.shape == aux["masks"][i].shape and torch.allclose(got, aux["masks"][i], atol=2e-6)
    inc = ref_utils.get_forward_masks(m, incremental=True)
    prev = torch.tensor(1.0)
    for i in range(4):
        exp = aux["masks"][i] * prev.ceil()
        assert torch.allclose(inc[f"encoder.layers.{i}"].cpu(), exp, atol=2e-6)
        prev = exp
    case = CASES["moevit"]
    sd, images = build_case(case)
    moe = build_model("vitmoe", case["cfg"])
    moe.load_state_dict(sd, strict=True)
    moe = moe.to("cuda:0").eval()
    moe.pk_precision = "fp32"
    moe(images.to("cuda:0"))
    gates = ref_utils.get_last_forward_gates(moe)
    _, aux = po.forward("moevit", sd, case["cfg"], images)
    assert sorted(gates) == ["encoder.layers.1.mlp", "encoder.layers.2.mlp"]
    for i in (1, 2):
        assert torch.equal(gates[f"encoder.layers.{i}.mlp"].cpu().argmax(-1), aux["mlp_gating"][i].argmax(-1))
