"""Drop-ins for reference models/vit.py."""
from .core import ViTBlock, ViTEncoder, VisionTransformer  # noqa: F401
