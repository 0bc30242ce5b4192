"""calciumgan_b200 — B200-native WGAN-GP training step behind CalciumGAN's plugin surface.

    from calciumgan_b200.models.registry import get_models
    from calciumgan_b200.algorithms.registry import get_algorithm

are drop-ins for the reference's gan.models.registry / gan.algorithms.registry.
"""
from .models.registry import get_models
from .algorithms.registry import get_algorithm

__all__ = ['get_models', 'get_algorithm']
