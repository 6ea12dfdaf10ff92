"""Model registry with the reference's aliases (reference models/models.py:15-86) for the five
families on the B200 hot path.  ``build_model`` keeps the reference signature; noise injection
and layer stitching are outside the hot-path scope and raise."""
from .core import (AdaptiveVisionTransformer, EEResidualVisionTransformer, RankVisionTransformer, ResidualVisionTransformer,
                   VisionTransformer, VisionTransformerMoE)

MODELS_MAP = {
    "visiontransformer": VisionTransformer, "VisionTransformer": VisionTransformer, "vit": VisionTransformer,
    "residualvisiontransformer": ResidualVisionTransformer, "ResidualVisionTransformer": ResidualVisionTransformer,
    "residualvit": ResidualVisionTransformer,
    "EEResidualVisionTransformer": EEResidualVisionTransformer, "eeResidualVisionTransformer": EEResidualVisionTransformer,
    "eeResidualvit": EEResidualVisionTransformer,
    "visiontransformermoe": VisionTransformerMoE, "VisionTransformerMoE": VisionTransformerMoE, "vitmoe": VisionTransformerMoE,
    "RankingVisionTransformer": RankVisionTransformer, "RankVisionTransformer": RankVisionTransformer,
    "AdaptiveVisionTransformer": AdaptiveVisionTransformer, "adavit": AdaptiveVisionTransformer,
}


def build_model(model_class, model_args, noise_args=None, remove_layers=None):
    if model_class not in MODELS_MAP:
        raise ValueError(f"Unknown model class {model_class}. Available models are {list(MODELS_MAP.keys())}")
    model_args = dict(model_args)
    # like the reference (models.py:69-73) pretrained-weight specs are dropped when rebuilding from a checkpoint
    model_args.pop("torch_pretrained_weights", None)
    model_args.pop("timm_pretrained_weights", None)
    model = MODELS_MAP[model_class](**model_args)
    if remove_layers is not None:
        model.remove_layers(list(remove_layers))
    if noise_args is not None and noise_args != {}:
        raise NotImplementedError("NoiseBlock injection (reference utils/utils.py:162-191) is outside the B200 hot-path scope")
    return model
