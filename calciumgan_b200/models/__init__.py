from .registry import register, get_models
from . import calciumgan
