from .registry import register, get_algorithm
from . import wgan_gp
