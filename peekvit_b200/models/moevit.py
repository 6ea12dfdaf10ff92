"""Drop-ins for reference models/moevit.py."""
from .core import MoE, TopKGate, MLPMoE, AttentionMoE, ViTBlockMoE, ViTEncoderMoE, VisionTransformerMoE  # noqa: F401
