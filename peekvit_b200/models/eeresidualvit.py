"""Drop-ins for reference models/eeresidualvit.py."""
from .core import EEResidualViTEncoder, EEResidualVisionTransformer  # noqa: F401
