"""Run the reference's OWN WGAN-GP code (unmodified files under /root/reference) over the torch-backed TensorFlow
stand-in in oracle/tf_shim -- the strongest pin this image allows for the oracle (TensorFlow 2.3.1 is not installable:
Python 3.12, no wheel, no network).

TEST INFRASTRUCTURE ONLY: used by tests/test_reference_shim.py and tests/golden/make_reference_golden.py. Nothing is copied
from the reference; its files are imported from REFERENCE_ROOT (default /root/reference), which exists in the build
container only -- the GPU box checks the committed fixture tests/golden/reference_step.npz instead.

What this pins: model construction (layer order, widths, kernel sizes, strides, LayerNorm / LeakyReLU placement,
PhaseShuffle after critic layers 1-4: calciumgan.py:22-192), Conv1DTranspose wrapping (models/utils.py:65-94), the
PhaseShuffle pad-and-slice logic incl. the draw order of the shifts (calciumgan.py:117-138), interpolation / gradient
penalty / losses / tape structure (wgan_gp.py:19-95), the update rule call order (optimizer.py:31-34) and the metrics
(gan.py:32-41, signals_metrics.py:9-28). What it cannot pin: TensorFlow's own kernels; their semantics are restated in
the shim (Keras defaults documented there) and, independently, in oracle/calciumgan_oracle.py.
"""
import argparse
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('REFERENCE_ROOT', '/root/reference')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tf_shim')


def available():
  return os.path.isfile(os.path.join(REFERENCE_ROOT, 'gan', 'algorithms', 'wgan_gp.py'))


def load_reference():
  """Import gan.models.calciumgan and gan.algorithms.wgan_gp from the reference tree. The packages' __init__ files import
  model families / algorithms that are not checked in (gan.models.conv1d, gan.algorithms.lswgan) and plotting /
  HDF5 libraries that are not installed, so the package objects are created empty with the reference directories as
  their __path__: every module that IS loaded is the reference's file, byte for byte."""
  if 'tensorflow' in sys.modules and not getattr(sys.modules['tensorflow'], '__file__', '').startswith(_SHIM):
    raise RuntimeError('a real tensorflow is already imported')
  if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)
  for name in ('h5py',):                     # imported at module top by gan/utils/h5_helper.py, unused on this path
    if name not in sys.modules:
      try:
        importlib.import_module(name)
      except ImportError:
        sys.modules[name] = types.ModuleType(name)
  for pkg, sub in (('gan', ''), ('gan.models', 'models'), ('gan.algorithms', 'algorithms'), ('gan.utils', 'utils')):
    if pkg not in sys.modules:
      m = types.ModuleType(pkg)
      m.__path__ = [os.path.join(REFERENCE_ROOT, 'gan', sub)]
      m.__package__ = pkg
      sys.modules[pkg] = m
  mods = types.SimpleNamespace()
  mods.tf = importlib.import_module('tensorflow')
  mods.signals_metrics = importlib.import_module('gan.utils.signals_metrics')
  setattr(sys.modules['gan.utils'], 'signals_metrics', mods.signals_metrics)
  mods.utils = importlib.import_module('gan.utils.utils')
  mods.calciumgan = importlib.import_module('gan.models.calciumgan')
  mods.wgan_gp = importlib.import_module('gan.algorithms.wgan_gp')
  for m in (mods.signals_metrics, mods.utils, mods.calciumgan, mods.wgan_gp):
    assert os.path.abspath(m.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), m.__file__
  return mods


def hparams_namespace(hp, batch_size):
  """The Namespace main.py + dataset_helper.py would hand to get_models / get_algorithm (oracle HParams -> reference)."""
  return argparse.Namespace(
      signal_shape=tuple(hp.signal_shape), num_channels=hp.num_channels, noise_dim=hp.noise_dim,
      noise_shape=(hp.noise_dim,), num_units=hp.num_units, kernel_size=hp.kernel_size, strides=hp.strides, m=hp.m,
      layer_norm=hp.layer_norm, batch_norm=False, activation='leakyrelu', normalize=hp.normalize,
      signals_min=hp.signals_min, signals_max=hp.signals_max, gradient_penalty=hp.gradient_penalty,
      n_critic=hp.n_critic, learning_rate=hp.learning_rate, conv2d=False, batch_size=batch_size,
      mixed_precision=False, model='calciumgan', algorithm='wgan-gp', verbose=0)


def build(hp, batch_size, gen_weights=None, dis_weights=None):
  """generator / discriminator / WGAN_GP objects built by the reference's own functions."""
  mods = load_reference()
  ns = hparams_namespace(hp, batch_size)
  generator, discriminator = mods.calciumgan.get_calciumgan(ns)        # calciumgan.py:10-12
  gan = mods.wgan_gp.WGAN_GP(ns, generator, discriminator, None)       # wgan_gp.py:12-17
  if gen_weights is not None:
    generator.set_weights(gen_weights)
  if dis_weights is not None:
    discriminator.set_weights(dis_weights)
  return mods, gan


def inject_step_randomness(tf, noises, alphas, shifts, n_critic):
  """Queue the draws in the order the reference consumes them (wgan_gp.py:64-95): per critic update noise, 4 shifts for
  D(real), 4 for D(fake), alpha, 4 for D(x_hat); then the generator update: noise, 4 shifts."""
  import numpy as np
  shifts = np.asarray(shifts).reshape(-1)
  assert len(shifts) == 12 * n_critic + 4
  tf.random.inject(normal=[noises[i] for i in range(n_critic + 1)], uniform=[alphas[i] for i in range(n_critic)],
                   ints=[int(s) for s in shifts])


def train_step(hp, gen_weights, dis_weights, real, noises, alphas, shifts):
  """One WGAN_GP.train(inputs) of the reference on injected weights / draws. Returns python floats + updated weights."""
  import torch
  mods, gan = build(hp, real.shape[0], gen_weights, dis_weights)
  inject_step_randomness(mods.tf, noises, alphas, shifts, hp.n_critic)
  gen_loss, dis_loss, gp, metrics = gan.train(torch.as_tensor(real, dtype=torch.float64))
  assert not mods.tf.random.normal_q and not mods.tf.random.uniform_q and not mods.tf.random.int_q, 'unused draws'
  return dict(gen_loss=float(gen_loss.detach()), dis_loss=float(dis_loss.detach()), gradient_penalty=float(gp.detach()),
              metrics={k: float(v.detach()) for k, v in metrics.items()},
              gen_weights=gan.generator.get_weights(), dis_weights=gan.discriminator.get_weights(),
              dis_grads=[[g.numpy() for g in gs] for gs in gan.dis_optimizer.optimizer.gradient_log],
              gen_grads=[[g.numpy() for g in gs] for gs in gan.gen_optimizer.optimizer.gradient_log],
              draw_order=list(mods.tf.random.log))


def sub_step_gradients(hp, gen_weights, dis_weights, real, noise_c, alpha, shifts_c, noise_g, shifts_g):
  """Per-parameter gradients as the reference's own _train_discriminator / _train_generator compute them on the GIVEN
  weights (wgan_gp.py:22-36,64-80 -> optimizer.py:31-34: tape.gradient w.r.t. model.trainable_variables), each from a
  fresh build so that neither sees the other's update."""
  import torch
  x = torch.as_tensor(real, dtype=torch.float64)
  mods, gan = build(hp, real.shape[0], gen_weights, dis_weights)
  mods.tf.random.inject(normal=[noise_c], uniform=[alpha], ints=[int(s) for s in shifts_c])
  dis_loss, gp = gan._train_discriminator(x)
  c = [g.numpy() for g in gan.dis_optimizer.optimizer.gradient_log[0]]
  mods, gan = build(hp, real.shape[0], gen_weights, dis_weights)
  mods.tf.random.inject(normal=[noise_g], uniform=[], ints=[int(s) for s in shifts_g])
  gen_loss, _ = gan._train_generator(x)
  g = [t.numpy() for t in gan.gen_optimizer.optimizer.gradient_log[0]]
  return dict(dis_loss=float(dis_loss.detach()), gradient_penalty=float(gp.detach()), gen_loss=float(gen_loss.detach()),
              dis_grads=c, gen_grads=g)
