"""peekvit_b200 — B200-native (sm_100a) encoder forward for alessiodevoto/peekvit.

Layout (only what the hot path needs):
  csrc/      hand-written CUDA kernels + the C ABI (include/peekvit_b200.h)
  _lib.py    ctypes binding of libpeekvit_b200.so
  ops.py     tensor-level wrappers (one C call each)
  engine.py  layer orchestration on packed token rows
  runner.py  model(images): validation, weight prepack cache, micro-batching
  models/    nn.Module mirrors of the reference classes (the drop-in boundary)
"""
from __future__ import annotations

import sys
import types

__version__ = "0.1.0"

_SHIM_MODULES = ("vit", "rankvit", "residualvit", "eeresidualvit", "adavit", "moevit", "models", "blocks", "adapters")


def install_as_peekvit() -> None:
    """Make ``peekvit.models.<x>.<Class>`` resolve to the B200 drop-ins, so the reference's
    hydra ``_target_`` paths (configs/model/*.yaml:1) and ``utils`` introspection helpers
    (``from peekvit.models.residualvit import ResidualModule``) keep working unchanged."""
    import importlib
    pkg = sys.modules.get("peekvit")
    if pkg is None:
        pkg = types.ModuleType("peekvit")
        pkg.__path__ = []
        sys.modules["peekvit"] = pkg
    models = importlib.import_module("peekvit_b200.models")
    sys.modules["peekvit.models"] = models
    pkg.models = models
    for name in _SHIM_MODULES:
        sys.modules[f"peekvit.models.{name}"] = importlib.import_module(f"peekvit_b200.models.{name}")
