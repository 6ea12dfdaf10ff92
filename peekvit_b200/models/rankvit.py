"""Drop-ins for reference models/rankvit.py."""
from .core import RankViTBlock, RankViTEncoder, RankVisionTransformer  # noqa: F401
