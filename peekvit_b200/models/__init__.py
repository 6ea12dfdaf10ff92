from .core import (VisionTransformer, RankVisionTransformer, ResidualVisionTransformer,  # noqa: F401
                   AdaptiveVisionTransformer, VisionTransformerMoE, EEResidualVisionTransformer)
from .models import MODELS_MAP, build_model  # noqa: F401
