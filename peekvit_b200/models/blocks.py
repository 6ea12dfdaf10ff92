"""Drop-ins for the parameter-bearing classes of reference models/blocks.py."""
from .core import MLP, SelfAttention  # noqa: F401
