"""Base algorithm class with the reference's surface (gan/algorithms/gan.py:10-97).

Only what WGAN-GP inherits is implemented (constructor, get_noise, metrics, validate,
generate); the vanilla BCE GAN losses of the reference's `'gan'` algorithm are out of scope.
"""
import torch

from .optimizer import Optimizer
from .. import _lib as L


METRIC_KEYS = ('signals_metrics/min', 'signals_metrics/max', 'signals_metrics/mean', 'signals_metrics/std')


def metrics_from_scalars(s):
  return {k: float(s[L.S_MET_MIN + i]) for i, k in enumerate(METRIC_KEYS)}


class GAN(object):

  def __init__(self, hparams, generator, discriminator, summary=None):
    if generator.engine is not discriminator.engine:
      raise ValueError('generator and discriminator must come from the same get_models() call')
    self.generator = generator
    self.discriminator = discriminator
    self.engine = generator.engine

    self._summary = summary
    self.noise_shape = tuple(hparams.noise_shape)
    self._normalize = hparams.normalize
    if hparams.normalize:
      self._signals_min = hparams.signals_min
      self._signals_max = hparams.signals_max

    self.gen_optimizer = Optimizer(hparams, self.engine, L.GENERATOR)
    self.dis_optimizer = Optimizer(hparams, self.engine, L.DISCRIMINATOR)

  def get_noise(self, batch_size):
    """gan.py:29-30."""
    return torch.randn((batch_size,) + self.noise_shape, device=self.engine.device)

  def generate(self, noise, denorm=False):
    """gan.py:92-97."""
    return self.engine.generate(noise, denorm=denorm)
