"""Drop-ins for reference models/residualvit.py."""
from .core import ResidualModule, ResidualGate, ResidualViTBlock, ResidualViTEncoder, ResidualVisionTransformer  # noqa: F401
