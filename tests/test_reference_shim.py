"""The oracle against the reference's OWN code.

TensorFlow 2.3.1 cannot be installed here, so the reference's files are executed unmodified from /root/reference over
a torch-backed stand-in for the ~40 TF / Keras calls they make (oracle/tf_shim, oracle/reference_runner.py). That pins the
oracle's structure -- layer order and widths, Conv1DTranspose wrapping, PhaseShuffle pad-and-slice incl. the order of the
random draws, interpolation, gradient penalty, losses, tape structure, Adam call order, metrics -- to the reference's
code rather than to a reading of it. The live tests need /root/reference (build container); the fixture test does not.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
from oracle import calciumgan_oracle as O
from oracle import reference_runner as R
from tests.util import rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_step.npz')
needs_reference = pytest.mark.skipif(not R.available(), reason='reference tree not present (GPU box)')
TOL = 1e-9   # both sides float64


def _oracle_step(hp, gw, dw, real, noises, alphas, shifts):
  st = O.TrainState.create(gw, dw)
  gen_loss, dis_loss, gp, metrics = O.train_step(st, real, noises, alphas, shifts, hp)
  return gen_loss, dis_loss, gp, metrics, [w.numpy() for w in st.gen], [w.numpy() for w in st.dis]


@needs_reference
def test_reference_train_step_paper_architecture():
  """The exact BASELINE.json architecture (2048 x 102, num_units 64, K 24, m 10), batch 1, one critic + one generator
  update, reference code vs oracle."""
  hp = O.HParams(n_critic=1)
  gw, dw = O.init_weights(hp, seed=21)
  real, noises, alphas, shifts = O.synthetic_batch(hp, 1, seed=22, n_critic=1)
  ref = R.train_step(hp, gw, dw, real, noises, alphas, shifts)
  gen_loss, dis_loss, gp, metrics, g_new, d_new = _oracle_step(hp, gw, dw, real, noises, alphas, shifts)
  for a, b in ((ref['gen_loss'], gen_loss), (ref['dis_loss'], dis_loss), (ref['gradient_penalty'], gp)):
    assert abs(a - b) <= TOL * max(1, abs(b))
  for a, b, w0 in list(zip(ref['gen_weights'], g_new, gw)) + list(zip(ref['dis_weights'], d_new, dw)):
    assert rel_err(a - w0, b - np.asarray(w0)) <= 1e-6


@needs_reference
@pytest.mark.parametrize('kw', [
    dict(signal_shape=(256, 20), noise_dim=8, num_units=16, kernel_size=24, m=3, n_critic=2),
    dict(signal_shape=(256, 12), noise_dim=4, num_units=8, kernel_size=24, m=10, n_critic=1),
    dict(signal_shape=(256, 6), noise_dim=4, num_units=4, kernel_size=5, m=2, n_critic=2, layer_norm=False),
    dict(signal_shape=(64, 10), noise_dim=4, num_units=4, kernel_size=6, m=1, n_critic=1, normalize=False),
])
def test_reference_train_step_matches_oracle(kw):
  hp = O.HParams(**kw)
  B = 3
  gw, dw = O.init_weights(hp, seed=11)
  gw, dw = O.randomize_weights(gw, 12), O.randomize_weights(dw, 13)
  real, noises, alphas, shifts = O.synthetic_batch(hp, B, seed=14)
  if any(k == 'm' and v == 10 for k, v in kw.items()):
    shifts = np.asarray(shifts).copy()
    shifts[:8] = [10, -10, 9, -9, -10, 10, 0, 1]          # both reflection branches at the extremes
  ref = R.train_step(hp, gw, dw, real, noises, alphas, shifts)
  gen_loss, dis_loss, gp, metrics, g_new, d_new = _oracle_step(hp, gw, dw, real, noises, alphas, shifts)
  assert abs(ref['gen_loss'] - gen_loss) <= TOL * max(1, abs(gen_loss))
  assert abs(ref['dis_loss'] - dis_loss) <= TOL * max(1, abs(dis_loss))
  assert abs(ref['gradient_penalty'] - gp) <= TOL * max(1, abs(gp))
  for k in metrics:
    assert abs(ref['metrics'][k] - metrics[k]) <= TOL * max(1, abs(metrics[k])), k
  assert len(ref['gen_weights']) == len(g_new) and len(ref['dis_weights']) == len(d_new)
  for a, b, w0 in list(zip(ref['gen_weights'], g_new, gw)) + list(zip(ref['dis_weights'], d_new, dw)):
    assert a.shape == b.shape
    assert rel_err(a - w0, b - np.asarray(w0)) <= 1e-7, rel_err(a - w0, b - np.asarray(w0))   # the Adam UPDATE, not the weight
  # the reference consumes its draws in the order the C ABI documents: noise, D(real) x4, D(fake) x4, alpha, D(x_hat) x4
  assert ref['draw_order'][:14] == ['normal(3, %d)' % hp.noise_dim] + ['int'] * 8 + ['uniform(3, 1, 1)'] + ['int'] * 4


@needs_reference
def test_reference_model_layout_matches_oracle():
  """Weight list order / shapes of the reference-built Keras models == the oracle's (checkpoint layout, utils.py:116-152)."""
  hp = O.HParams()
  mods, gan = R.build(hp, 2)
  gw, dw = O.init_weights(hp, seed=0)
  assert [tuple(w.shape) for w in gan.generator.get_weights()] == [tuple(np.shape(w)) for w in gw]
  assert [tuple(w.shape) for w in gan.discriminator.get_weights()] == [tuple(np.shape(w)) for w in dw]
  assert sum(int(np.prod(w.shape)) for w in gan.generator.get_weights()) == 4375740
  assert sum(int(np.prod(w.shape)) for w in gan.discriminator.get_weights()) == 4110273
  with pytest.raises(ValueError):          # calciumgan.py:17-18: length not divisible by strides^5
    R.build(O.HParams(signal_shape=(100, 4)), 2)


@needs_reference
def test_reference_phase_shuffle_matches_closed_form():
  """PhaseShuffle.call (calciumgan.py:117-138) executed from the reference file vs the oracle's closed-form index map,
  every shift in [-m, m]; bit-exact (pure data movement)."""
  mods = R.load_reference()
  w, m = 16, 10
  x = np.random.RandomState(0).standard_normal((2, w, 3))
  layer = mods.calciumgan.PhaseShuffle((None, w, 3), m=m)
  for shift in range(-m, m + 1):
    mods.tf.random.inject(ints=[shift])
    got = layer(torch.as_tensor(x, dtype=torch.float64)).numpy()
    np.testing.assert_array_equal(got, x[:, O.phase_shuffle_index(w, shift), :])


def test_oracle_matches_reference_fixture():
  """tests/golden/reference_step.npz was produced by the reference's code (make_reference_golden.py); the oracle must
  reproduce it. Runs everywhere (no reference tree needed)."""
  import make_reference_golden as G
  gold = np.load(GOLD)
  hp, gw, dw, real, noises, alphas, shifts = G.inputs()
  fake, gen_loss, dis_loss, gp, metrics = O.validate_step(gw, dw, real, noises[0], alphas[0],
                                                          np.asarray(shifts[:12]).reshape(3, 4), hp)
  assert rel_err(fake.numpy(), gold['val_fake']) <= TOL
  got = np.array([gen_loss, dis_loss, gp] + [metrics[k] for k in sorted(metrics)])
  assert rel_err(got, gold['val_scalars']) <= TOL
  gen_loss, dis_loss, gp, metrics, g_new, d_new = _oracle_step(hp, gw, dw, real, noises, alphas, shifts)
  got = np.array([gen_loss, dis_loss, gp] + [metrics[k] for k in sorted(metrics)])
  assert rel_err(got, gold['train_scalars']) <= TOL
  for i, (w, w0) in enumerate(zip(g_new, gw)):
    assert rel_err(w - np.asarray(w0), gold['gen_w_%02d' % i] - np.asarray(w0)) <= 1e-7, ('gen', i)
  for i, (w, w0) in enumerate(zip(d_new, dw)):
    assert rel_err(w - np.asarray(w0), gold['dis_w_%02d' % i] - np.asarray(w0)) <= 1e-7, ('dis', i)


def test_reference_sub_step_gradients_equal_the_oracle():
  """The gradients the reference's own _train_discriminator / _train_generator hand to Adam (optimizer.py:31-34),
  captured from the stand-in optimizer, equal the oracle's per-parameter gradients; the committed fixtures carry them
  to the GPU box."""
  import sys
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
  import make_reference_golden as G
  hp, gw, dw, real, noises, alphas, shifts = G.inputs()
  sg = R.sub_step_gradients(hp, gw, dw, real, noises[0], alphas[0], shifts[:12], noises[1], shifts[12:16])
  c = O.critic_step(gw, dw, real, noises[0], alphas[0], shifts[:12].reshape(3, 4), hp)
  g = O.generator_step(gw, dw, real, noises[1], shifts[12:16], hp)
  assert len(sg['dis_grads']) == 12 and len(sg['gen_grads']) == 24
  for a, b in zip(sg['dis_grads'], c['grads']):
    assert np.abs(a - b.numpy()).max() <= 1e-9 * max(1.0, np.abs(a).max())
  for a, b in zip(sg['gen_grads'], g['grads']):
    assert np.abs(a - b.numpy()).max() <= 1e-9 * max(1.0, np.abs(a).max())
  gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_step.npz'))
  for i, a in enumerate(sg['dis_grads']):
    np.testing.assert_allclose(gold['c_grad_%02d' % i], a, rtol=1e-6, atol=1e-12)
