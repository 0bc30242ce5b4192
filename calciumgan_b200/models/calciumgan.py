"""`calciumgan` model plugin: generator / critic handles over one CUDA engine.

Mirrors gan/models/calciumgan.py:10-19,22-103,141-192 of the reference: same builder name,
same hyper-parameters (signal_shape, noise_dim, num_units, kernel_size, strides, m,
layer_norm, normalize), same get_weights()/set_weights() order and layouts. The Keras graph
is replaced by libcalciumgan_b200.so; both handles share one device context.
"""
import numpy as np

from .registry import register
from .. import _lib as L
from ..engine import Engine, hparams_to_config


def calculate_noise_shape(output_shape, noise_dim, num_convolutions, strides):
  """gan/models/calciumgan.py:15-19."""
  w = output_shape[0] / (strides**num_convolutions)
  if not float(w).is_integer():
    raise ValueError('Conv1D: w {} is not an integer.'.format(w))
  return (int(w), noise_dim)


class Variable(object):
  """Minimal stand-in for tf.Variable as read by summary_helper.py:544-557 (.name, .numpy())."""

  def __init__(self, model, index, name, shape):
    self._model, self._index, self.name, self.shape = model, index, name, tuple(shape)

  def numpy(self):
    return self._model.get_weights()[self._index]

  def __array__(self, dtype=None):
    a = self.numpy()
    return a.astype(dtype) if dtype is not None else a


class ModelHandle(object):
  """What the reference calls `generator` / `discriminator` (a tf.keras.Model)."""

  def __init__(self, engine, which, name, layer_names):
    self.engine, self.which, self.name = engine, which, name
    infos = engine.tensor_infos(which)
    assert len(infos) == len(layer_names)
    self.trainable_variables = [
        Variable(self, i, n, s) for i, (n, (s, _)) in enumerate(zip(layer_names, infos))]

  @property
  def trainable_weights(self):
    return self.trainable_variables

  def get_weights(self):
    return self.engine.get_weights(self.which)

  def set_weights(self, weights):
    self.engine.set_weights(self.which, weights)

  def count_params(self):
    return self.engine.num_params(self.which)

  def summary(self):
    print('Model: "{}"'.format(self.name))
    for v in self.trainable_variables:
      print('  {:40s} {}'.format(v.name, v.shape))
    print('Trainable params: {:,}'.format(self.count_params()))

  def __call__(self, inputs, training=False, shifts=None):
    if self.which == L.GENERATOR:
      return self.engine.generate(inputs, denorm=False)
    if shifts is None:   # PhaseShuffle is active regardless of `training` (calciumgan.py:117)
      m = self.engine.cfg.phase_m
      shifts = np.random.randint(-m, m + 1, size=4)
    return self.engine.critic_forward(inputs, shifts)


def _generator_names(layer_norm):
  names = ['dense/kernel:0', 'dense/bias:0']
  for i in range(5):
    p = 'conv1d_transpose%s/conv2d_transpose%s/' % (('_%d' % i) if i else '', ('_%d' % i) if i else '')
    names += [p + 'kernel:0', p + 'bias:0']
    if layer_norm:
      q = 'layer_normalization%s/' % (('_%d' % i) if i else '')
      names += [q + 'gamma:0', q + 'beta:0']
  return names + ['dense_1/kernel:0', 'dense_1/bias:0']


def _discriminator_names():
  names = []
  for i in range(5):
    p = 'conv1d%s/' % (('_%d' % i) if i else '')
    names += [p + 'kernel:0', p + 'bias:0']
  return names + ['dense_2/kernel:0', 'dense_2/bias:0']


def build_engine(hparams):
  """One engine per hparams Namespace; world_size/rank come from torch.distributed if initialised."""
  import torch.distributed as dist
  world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_available() and dist.is_initialized() else (1, 0)
  calculate_noise_shape(hparams.signal_shape, hparams.noise_dim, 5, hparams.strides)
  cfg = hparams_to_config(hparams, world_size=world, rank=rank,
                          force_simt=getattr(hparams, 'force_simt', False),
                          debug_flags=getattr(hparams, 'debug_flags', 0))
  engine = Engine(cfg)
  seed = int(getattr(hparams, 'seed', 1234))   # main.py:11-12
  engine.init_weights(seed)
  engine.seed(seed)
  return engine


def generator(hparams, engine=None):
  engine = engine or build_engine(hparams)
  return ModelHandle(engine, L.GENERATOR, 'generator', _generator_names(bool(hparams.layer_norm)))


def discriminator(hparams, engine=None):
  engine = engine or build_engine(hparams)
  return ModelHandle(engine, L.DISCRIMINATOR, 'discriminator', _discriminator_names())


@register('calciumgan')
def get_calciumgan(hparams):
  engine = build_engine(hparams)
  return generator(hparams, engine), discriminator(hparams, engine)
