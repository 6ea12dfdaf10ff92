"""CPU oracle for the peekvit encoder-forward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline — never as the thing shipped.  The product path
(``peekvit_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md §4, §8c), so the oracle is pinned against *outputs of the reference
itself run in the authoring container*: ``tests/golden/make_golden.py`` imports
``/root/reference`` as package ``peekvit``, loads the oracle's seeded state dict
into the reference modules (``strict=True``), runs the reference forward and
stores logits / masks / kept-token indices / halting state as small ``.npz``
fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py`` replays them.
"""
from .weights import make_state_dict, synthetic_images, FAMILIES  # noqa: F401
from .peekvit_oracle import (  # noqa: F401
    layer_norm, mha, mlp, patch_embed, vit_block, stable_topk_desc,
    vit_forward, rankvit_forward, residualvit_forward, avit_forward, moevit_forward, eeresidualvit_forward,
    forward, flops_per_image,
)
