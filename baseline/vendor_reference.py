"""Vendor the UNMODIFIED reference tree for the benchmark's reference arm.

    python baseline/vendor_reference.py            # /root/reference -> baseline/_ref/peekvit

The reference has no setup.py / pyproject.toml (only requirements.txt), so ``pip install --target`` has nothing to
build; its entry scripts import it as the package ``peekvit`` = the repository root (utils/utils.py:2, validate/test.py:2,
train/train.py:3).  This script therefore copies the tree as it is (sources only: no notebooks, images or git metadata)
to ``baseline/_ref/peekvit``.  ``baseline/_ref/`` is git-ignored (no reference source enters the history) but is NOT
gpurun-ignored, so it travels to the GPU box with the snapshot like the built ``.so``.  Nothing in ``peekvit_b200``
imports it; ``bench.py --impl reference`` and the ``gpu_eager_reference`` / ``cpu_baseline`` legs do.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PEEKVIT_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref", "peekvit")


def vendor(force: bool = False) -> str:
    if not os.path.isdir(SRC):
        if os.path.isdir(DST):
            return DST
        raise FileNotFoundError(f"{SRC} not found and {DST} was not vendored earlier")
    if os.path.isdir(DST):
        if not force:
            return DST
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns(".git", "images", "notebooks", "__pycache__", "*.ipynb", "*.png", "*.jpg"))
    return DST        # no __init__.py is added: `peekvit` is a namespace package, exactly like the reference's own layout


def import_reference():
    """Put ``baseline/_ref`` on sys.path and return the reference's model classes (KeyError-free dict)."""
    root = os.path.join(HERE, "_ref")
    if not os.path.isdir(os.path.join(root, "peekvit", "models")):
        raise ImportError("baseline/_ref/peekvit is absent: run `python baseline/vendor_reference.py` where /root/reference exists")
    if root not in sys.path:
        sys.path.insert(0, root)
    from peekvit.models.vit import VisionTransformer
    from peekvit.models.rankvit import RankVisionTransformer
    from peekvit.models.residualvit import ResidualVisionTransformer
    from peekvit.models.adavit import AdaptiveVisionTransformer
    from peekvit.models.moevit import VisionTransformerMoE
    return {"vit": VisionTransformer, "rankvit": RankVisionTransformer, "residualvit": ResidualVisionTransformer,
            "adavit": AdaptiveVisionTransformer, "moevit": VisionTransformerMoE}


if __name__ == "__main__":
    print("vendored to", vendor(force=True))
