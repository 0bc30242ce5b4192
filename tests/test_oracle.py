"""Self-consistency of the CPU oracle (the reference has no golden vectors: parity unpinned).

(i) torch convs == definitional naive loops (Keras SAME arithmetic),
(ii) PhaseShuffle closed form == literal pad+slice transcription (calciumgan.py:126-137),
(iii) autograd gradient penalty == hand-derived 4-pass GP,
(iv) finite differences of the critic loss,
(v) committed golden fixture (tests/golden) regenerates bit-for-bit.
"""
import os

import numpy as np
import pytest
import torch

from oracle import calciumgan_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('K', [24, 6, 5, 2])
def test_conv_matches_naive(K):
  rng = np.random.RandomState(0)
  x = rng.standard_normal((2, 32, 3))
  w = rng.standard_normal((K, 3, 5))
  b = rng.standard_normal(5)
  y = O.conv1d_same(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
  np.testing.assert_allclose(y, O.naive_conv1d_same(x, w, b), atol=1e-12)


@pytest.mark.parametrize('K', [24, 6, 5, 2])
def test_conv_transpose_matches_naive(K):
  rng = np.random.RandomState(1)
  x = rng.standard_normal((2, 16, 3))
  w = rng.standard_normal((K, 1, 5, 3))
  b = rng.standard_normal(5)
  y = O.conv1d_transpose_same(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
  assert y.shape == (2, 32, 5)
  np.testing.assert_allclose(y, O.naive_conv1d_transpose_same(x, w, b), atol=1e-12)


@pytest.mark.parametrize('w', [4, 11, 64, 128])
def test_phase_shuffle_closed_form_is_literal(w):
  x = np.arange(2 * w * 3, dtype=np.float32).reshape(2, w, 3)
  m = min(10, w - 1)
  for shift in range(-m, m + 1):
    lit = O.phase_shuffle_literal(x, shift)
    idx = O.phase_shuffle_index(w, shift)
    assert idx.dtype == np.int32 and idx.min() >= 0 and idx.max() < w
    np.testing.assert_array_equal(lit, x[:, idx, :])
    np.testing.assert_array_equal(O.phase_shuffle(torch.tensor(x), shift).numpy(), lit)


def test_dgrad_wgrad_match_autograd():
  rng = np.random.RandomState(2)
  x = torch.tensor(rng.standard_normal((2, 32, 3)), requires_grad=True)
  w = torch.tensor(rng.standard_normal((6, 3, 5)), requires_grad=True)
  dy = torch.tensor(rng.standard_normal((2, 16, 5)))
  y = O.conv1d_same(x, w, None)
  gx, gw = torch.autograd.grad((y * dy).sum(), [x, w])
  np.testing.assert_allclose(O.conv1d_same_dgrad(dy, w.detach(), 32).numpy(), gx.numpy(), atol=1e-12)
  np.testing.assert_allclose(O.conv1d_same_wgrad(x.detach(), dy, 6).numpy(), gw.numpy(), atol=1e-12)


def test_gp_four_pass_matches_autograd(tiny_hp):
  hp = tiny_hp
  gw, dw = O.init_weights(hp, seed=3)
  dw = O.randomize_weights(dw, 4)
  real, noises, alphas, shifts = O.synthetic_batch(hp, 3, seed=5)
  xhat = real * 0.3 + 0.1
  sh = [2, -1, 0, -2]
  dwt = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in dw]
  x = torch.tensor(xhat, dtype=torch.float64, requires_grad=True)
  out = O.discriminator_forward(dwt, x, sh, hp)
  (g,) = torch.autograd.grad(out.sum(), x, create_graph=True)
  gp = ((torch.sqrt((g.reshape(3, -1)**2).sum(1)) - 1)**2).mean()
  ref = torch.autograd.grad(gp, dwt, allow_unused=True)
  gp2, g2, grads = O.gp_four_pass(dw, xhat, sh, hp)
  assert abs(float(gp.detach()) - float(gp2)) < 1e-12
  np.testing.assert_allclose(g2.numpy(), g.detach().numpy(), atol=1e-12)
  for i, (a, b) in enumerate(zip(ref, grads)):
    if a is None:   # biases: dGP/db = 0 exactly
      assert i % 2 == 1 and float(b.abs().max()) == 0.0
      continue
    np.testing.assert_allclose(b.numpy(), a.numpy(), rtol=1e-9, atol=1e-11)


def test_critic_loss_finite_difference(tiny_hp):
  hp = tiny_hp
  gw, dw = O.init_weights(hp, seed=7)
  gw, dw = O.randomize_weights(gw, 8), O.randomize_weights(dw, 9)
  real, noises, alphas, shifts = O.synthetic_batch(hp, 2, seed=10)
  sh = shifts[:12].reshape(3, 4)
  r = O.critic_step(gw, dw, real, noises[0], alphas[0], sh, hp)
  rng = np.random.RandomState(11)
  for ti in (0, 1, 4, 10):
    d = rng.standard_normal(dw[ti].shape)
    eps = 1e-6
    wp = [a.astype(np.float64) for a in dw]
    wm = [a.astype(np.float64) for a in dw]
    wp[ti] = wp[ti] + eps * d
    wm[ti] = wm[ti] - eps * d
    lp = O.critic_step(gw, wp, real, noises[0], alphas[0], sh, hp)['dis_loss']
    lm = O.critic_step(gw, wm, real, noises[0], alphas[0], sh, hp)['dis_loss']
    fd = (lp - lm) / (2 * eps)
    an = float((r['grads'][ti].numpy() * d).sum())
    assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)), (ti, fd, an)


def test_weight_shapes_paper_config():
  hp = O.HParams()
  gs, ds = O.weight_shapes(hp)
  assert len(gs) == 24 and len(ds) == 12
  assert sum(int(np.prod(s)) for s in gs) == 4375740
  assert sum(int(np.prod(s)) for s in ds) == 4110273
  with pytest.raises(ValueError):
    O.calculate_noise_shape((2050, 102), 32, 5, 2)


def test_adam_is_keras_form():
  w, m, v = torch.tensor([1.0], dtype=torch.float64), torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64)
  g = torch.tensor([0.5], dtype=torch.float64)
  w1, m1, v1 = O.adam_update(w, m, v, g, 1, 1e-4)
  lr_t = 1e-4 * np.sqrt(1 - 0.999) / (1 - 0.9)
  exp = 1.0 - lr_t * 0.05 / (np.sqrt(0.001 * 0.25) + 1e-7)
  assert abs(float(w1) - exp) < 1e-9


def test_golden_fixture_regenerates():
  import tests.golden.make_golden as mk
  path = os.path.join(GOLD, 'tiny_step.npz')
  assert os.path.exists(path), 'run python tests/golden/make_golden.py'
  gold = np.load(path)
  fresh = mk.compute()
  assert set(gold.files) == set(fresh.keys())
  for k in gold.files:
    np.testing.assert_allclose(fresh[k], gold[k], rtol=1e-10, atol=1e-12, err_msg=k)
