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

  def metrics(self, real, fake):
    """gan.py:32-41: signals_metrics/{min,max,mean,std} of (real, fake), de-normalised when hparams.normalize."""
    return dict(zip(METRIC_KEYS, self.engine.metrics(real, fake)))

  def _step(self, real, noise, training=True, alpha=None, shifts=None):
    """gan.py:58-70 -> (fake, gen_loss, dis_loss, gradient_penalty, metrics), with the losses of the subclass
    (WGAN-GP: wgan_gp.py:19-20,52-62). PhaseShuffle is active whatever `training` says (calciumgan.py:117) and the
    models have no other train / inference difference (no dropout, no batch norm), so both values run the same
    kernels; nothing is updated."""
    fake, s = self.engine.validate(real, noise, alpha, shifts)
    return (fake, float(s[L.S_GEN_LOSS]), float(s[L.S_DIS_LOSS]), float(s[L.S_GP]), metrics_from_scalars(s))

  def generate(self, noise, denorm=False):
    """gan.py:92-97."""
    return self.engine.generate(noise, denorm=denorm)
