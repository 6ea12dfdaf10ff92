"""Drop-ins for reference models/adavit.py."""
from .core import AViTBlock, AViTEncoder, AdaptiveVisionTransformer  # noqa: F401
