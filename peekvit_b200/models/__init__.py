from .core import (VisionTransformer, RankVisionTransformer, ResidualVisionTransformer,  # noqa: F401
                   AdaptiveVisionTransformer, VisionTransformerMoE, EEResidualVisionTransformer)
from .models import MODELS_MAP, add_noise, build_model  # noqa: F401
from .core import NoiseBlock  # noqa: F401
