"""Model registry with the reference's aliases (reference models/models.py:15-86) for the families on the B200 hot
path.  ``build_model`` keeps the reference signature, including ``noise_args`` (a NoiseBlock spliced into
``encoder.layers``, utils/utils.py:162-191) and ``remove_layers``."""
from collections import OrderedDict

import torch

from .core import NoiseBlock
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
        noise_module = add_noise(model, **noise_args)
        noise_module.set_value(0.0)                 # models.py:80-83: loaded with the noise switched off
        print("Loaded model with noise. Noise will be set to 0.0, you can change this by calling "
              "model.noise_module.set_value(new_noise_value)")
    return model


def add_noise(model, layer: int, noise_type: str, std: float = None, snr: float = None, prob: float = None, **kwargs):
    """Mirror of reference utils/utils.py:162-191: insert a NoiseBlock before ``encoder.layers[layer]`` and return it."""
    noise_module = NoiseBlock(noise_type=noise_type, std=std, snr=snr, prob=prob)
    new_layers = list(model.encoder.layers)
    if new_layers and isinstance(new_layers[0], tuple):
        new_layers.insert(layer, ("noise", noise_module))
        model.encoder.layers = torch.nn.Sequential(OrderedDict(new_layers))
    else:
        new_layers.insert(layer, noise_module)
        model.encoder.layers = torch.nn.Sequential(*new_layers)
    return noise_module
