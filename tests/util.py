import argparse

import numpy as np


def namespace_from_oracle(hp, batch_size, mixed_precision=False, **kw):
  """hparams Namespace as main.py + dataset_helper.py would build it."""
  ns = argparse.Namespace(
      signal_shape=tuple(hp.signal_shape), num_channels=hp.num_channels, noise_dim=hp.noise_dim,
      noise_shape=(hp.noise_dim,), num_units=hp.num_units, kernel_size=hp.kernel_size, strides=hp.strides,
      m=hp.m, layer_norm=hp.layer_norm, batch_norm=False, activation='leakyrelu', normalize=hp.normalize,
      signals_min=hp.signals_min, signals_max=hp.signals_max, gradient_penalty=hp.gradient_penalty,
      n_critic=hp.n_critic, learning_rate=hp.learning_rate, conv2d=False, batch_size=batch_size,
      mixed_precision=mixed_precision, model='calciumgan', algorithm='wgan-gp', verbose=0)
  for k, v in kw.items():
    setattr(ns, k, v)
  return ns


def rel_err(a, b):
  """norm-wise relative error ||a-b|| / ||b|| (0 if both are zero)."""
  a = np.asarray(a, np.float64)
  b = np.asarray(b, np.float64)
  d = np.linalg.norm((a - b).ravel())
  n = np.linalg.norm(b.ravel())
  return 0.0 if d == 0 else d / max(n, 1e-30)
