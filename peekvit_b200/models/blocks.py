"""Drop-ins for the classes of reference models/blocks.py that the model families and utils.add_noise use."""
from .core import MLP, SelfAttention, NoiseBlock  # noqa: F401
