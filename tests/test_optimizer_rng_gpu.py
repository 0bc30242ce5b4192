"""cg_apply_update (fused Adam + refresh of the packed bf16 GEMM operands) and the library's random streams, tested
directly through the C ABI.

Adam: Keras form of gan/algorithms/optimizer.py:9,34 -- lr_t = lr sqrt(1 - b2^t) / (1 - b1^t), w -= lr_t m / (sqrt(v) + 1e-7)
against oracle.adam_update with non-zero moments at t = 1, 7, 1000 and data-parallel world sizes 1 and 8 (gradient
scaled by 1 / world_size); a non-finite gradient skips the update (what the reference's LossScaleOptimizer does,
optimizer.py:10-12). Random streams: noise ~ N(0, 1) (gan.py:29-30), alpha ~ U[0, 1) (wgan_gp.py:40), PhaseShuffle shifts
uniform on {-m..m} (calciumgan.py:121-124), rank-specific noise / alpha, rank-SHARED shifts whatever a rank injects."""
import numpy as np
import pytest
import torch

from oracle import calciumgan_oracle as O
from calciumgan_b200 import _lib as L
from tests.util import namespace_from_oracle, rel_err

pytestmark = pytest.mark.gpu


def _hp(**kw):
  d = dict(signal_shape=(256, 20), noise_dim=8, num_units=16, kernel_size=24, m=3, n_critic=2)
  d.update(kw)
  return O.HParams(**d)


def engine(hp, batch, mixed=False, world_size=1, rank=0, debug_flags=0, seed=1234):
  from calciumgan_b200.engine import Engine, hparams_to_config
  ns = namespace_from_oracle(hp, batch, mixed_precision=mixed)
  eng = Engine(hparams_to_config(ns, world_size=world_size, rank=rank, debug_flags=debug_flags))
  eng.init_weights(seed)
  eng.seed(seed)
  return eng


@pytest.mark.parametrize('which', [L.GENERATOR, L.DISCRIMINATOR])
@pytest.mark.parametrize('world_size', [1, 8])
@pytest.mark.parametrize('t', [1, 7, 1000])
def test_adam_kernel_against_oracle(which, world_size, t):
  hp = _hp()
  eng = engine(hp, 2, mixed=True, world_size=world_size)
  rng = np.random.RandomState(100 * t + world_size + which)
  w0 = eng.get_weights(which)
  w = [(a + 0.01 * rng.standard_normal(a.shape)).astype(np.float32) for a in w0]
  m = [(1e-2 * rng.standard_normal(a.shape)).astype(np.float32) for a in w0]
  v = [(1e-4 * rng.uniform(0.0, 1.0, a.shape)).astype(np.float32) for a in w0]
  g = [(rng.standard_normal(a.shape) * 10.0 ** rng.uniform(-4, 0)).astype(np.float32) for a in w0]
  flat = lambda xs: np.concatenate([x.ravel() for x in xs])
  eng.set_weights(which, w)
  eng.set_opt_state(which, flat(m), flat(v), step=t - 1)
  eng.set_grads(which, g)
  eng.apply_update(which)
  got_w = flat(eng.get_weights(which))
  got_m, got_v, step = eng.get_opt_state(which)
  assert step == t and eng.skipped_updates(which) == 0
  T = lambda x: torch.tensor(flat(x), dtype=torch.float64)
  # Keras holds beta_1 / beta_2 as float32 tensors (0.999f = 0.99900001287...): same constants in the oracle
  rw, rm, rv = O.adam_update(T(w), T(m), T(v), T(g) / world_size, t, hp.learning_rate, b1=float(np.float32(0.9)),
                             b2=float(np.float32(0.999)))
  assert rel_err(got_m, rm.numpy()) <= 1e-6
  assert rel_err(got_v, rv.numpy()) <= 1e-6
  assert rel_err(got_w, rw.numpy()) <= 1e-6
  # the update itself (|dw| ~ lr): fp32 sqrt / divide / one rounding of w
  dw_ref = rw.numpy() - flat(w).astype(np.float64)
  dw_got = got_w.astype(np.float64) - flat(w).astype(np.float64)
  assert np.abs(dw_got - dw_ref).max() <= 1e-3 * np.abs(dw_ref).max() + 6e-8 * np.abs(flat(w)).max()
  eng.close()


def test_fused_adam_matches_two_kernel_form_and_refreshes_the_packed_weights():
  """adam_pack_kernel vs adam_kernel + pack_weights_kernel (CG_DEBUG_NO_ADAM_FUSE): identical master weights and
  moments, and identical packed bf16 operands -- a critic / generator forward after the update is bit-identical to the
  one of a fresh engine that loaded the updated weights through set_weights (full re-pack)."""
  hp = _hp(signal_shape=(512, 102), num_units=32)
  B = 3
  real, noises, alphas, shifts = O.synthetic_batch(hp, B, seed=5, n_critic=2)
  engs = [engine(hp, B, mixed=True, debug_flags=f) for f in (0, L.DEBUG_NO_ADAM_FUSE)]
  rng = np.random.RandomState(9)
  shapes = {w: [a.shape for a in engs[0].get_weights(w)] for w in (L.GENERATOR, L.DISCRIMINATOR)}
  # identical gradients for both engines (the step's own weight-gradient kernels accumulate with atomics, so two runs
  # of a step are not bit-reproducible; Adam and the re-pack are)
  grads = [{w: [(1e-2 * rng.standard_normal(s)).astype(np.float32) for s in shapes[w]] for w in shapes} for _ in range(2)]
  outs = []
  for eng in engs:
    for it in range(2):
      for w in shapes:
        eng.set_grads(w, grads[it][w])
        eng.apply_update(w)
    outs.append((eng.get_weights(L.GENERATOR), eng.get_weights(L.DISCRIMINATOR), eng.get_opt_state(L.GENERATOR),
                 eng.get_opt_state(L.DISCRIMINATOR)))
  for a, b in zip(outs[0][0] + outs[0][1], outs[1][0] + outs[1][1]):
    np.testing.assert_array_equal(a, b)
  for k in (2, 3):
    np.testing.assert_array_equal(outs[0][k][0], outs[1][k][0])
    np.testing.assert_array_equal(outs[0][k][1], outs[1][k][1])
    assert outs[0][k][2] == outs[1][k][2] == 2
  fresh = engine(hp, B, mixed=True)
  fresh.set_weights(L.GENERATOR, outs[0][0])
  fresh.set_weights(L.DISCRIMINATOR, outs[0][1])
  sh4 = shifts[:4]
  for eng in engs:
    assert torch.equal(eng.critic_forward(real, sh4), fresh.critic_forward(real, sh4))
    assert torch.equal(eng.generate(noises[0]), fresh.generate(noises[0]))
  fp = engine(hp, B, mixed=False)     # the fp32 path packs float copies through the same kernel
  fp.critic_step(real, noises[0], alphas[0], shifts[:12])
  fresh32 = engine(hp, B, mixed=False)
  fresh32.set_weights(L.DISCRIMINATOR, fp.get_weights(L.DISCRIMINATOR))
  assert torch.equal(fp.critic_forward(real, sh4), fresh32.critic_forward(real, sh4))


@pytest.mark.parametrize('bad', [np.nan, np.inf])
def test_non_finite_gradient_skips_the_update(bad):
  hp = _hp()
  eng = engine(hp, 2, mixed=True)
  for which in (L.GENERATOR, L.DISCRIMINATOR):
    w0 = eng.get_weights(which)
    g = [np.full(a.shape, 1e-3, np.float32) for a in w0]
    g[-2].reshape(-1)[-1] = bad
    eng.set_grads(which, g)
    eng.apply_update(which)
    for a, b in zip(w0, eng.get_weights(which)):
      np.testing.assert_array_equal(a, b)
    m, v, step = eng.get_opt_state(which)
    assert step == 0 and not m.any() and not v.any() and eng.skipped_updates(which) == 1
    g[-2].reshape(-1)[-1] = 1e-3
    eng.set_grads(which, g)
    eng.apply_update(which)
    assert eng.get_step(which) == 1 and eng.skipped_updates(which) == 1
    assert all(np.isfinite(a).all() for a in eng.get_weights(which))
    assert not np.array_equal(w0[0], eng.get_weights(which)[0])


def test_library_drawn_noise_alpha_and_shifts():
  from scipy import stats
  hp = _hp(m=10, signal_shape=(2048, 20), n_critic=5)
  B = 64
  eng = engine(hp, B, mixed=True, seed=77)
  real = np.random.RandomState(0).uniform(0, 1, size=(B, 2048, 20)).astype(np.float32)
  nc, nd = hp.n_critic, hp.noise_dim
  all_shifts, noise_sets = [], []
  for _ in range(40):
    eng.train_step(real)
    noise, alpha, sh = eng.last_draws((nc + 1) * B * nd, nc * B, 12 * nc + 4)
    all_shifts.append(sh.copy())
    noise_sets.append(noise.cpu().numpy().astype(np.float64))
    a = alpha.cpu().numpy()
    assert a.min() >= 0.0 and a.max() < 1.0
  noise = np.concatenate(noise_sets)        # 40 x 6 x 64 x 8 = 122 880 values
  assert abs(noise.mean()) <= 4.0 / np.sqrt(noise.size) and abs(noise.var() - 1.0) <= 4.0 * np.sqrt(2.0 / noise.size)
  assert abs(stats.skew(noise)) <= 0.05 and abs(stats.kurtosis(noise)) <= 0.1
  assert stats.kstest(noise, 'norm').pvalue > 1e-3
  assert len(np.unique(noise)) > 0.99 * noise.size          # sub-steps do not repeat each other's noise
  assert stats.kstest(a.astype(np.float64), 'uniform').pvalue > 1e-3
  sh = np.concatenate(all_shifts)
  assert sh.min() == -10 and sh.max() == 10
  counts = np.bincount(sh + 10, minlength=21)
  assert (counts > 0).all() and stats.chisquare(counts).pvalue > 1e-3, counts
  # reproducible after cg_seed
  eng.seed(77)
  eng.train_step(real)
  n2, a2, s2 = eng.last_draws((nc + 1) * B * nd, nc * B, 12 * nc + 4)
  np.testing.assert_array_equal(s2, all_shifts[0])
  np.testing.assert_array_equal(n2.cpu().numpy().astype(np.float64), noise_sets[0])
  eng.seed(78)
  eng.train_step(real)
  assert not np.array_equal(eng.last_draws(0, 0, 12 * nc + 4)[2], all_shifts[0])


def test_ranks_share_shifts_but_not_noise_whatever_they_inject():
  """Data parallelism: PhaseShuffle shifts are per-call scalars shared by the global batch, so every rank
  must draw the same ones; noise and alpha are per-sample. A rank that injects its noise / alpha (or a rank that runs
  the sub-steps one by one, as the data-parallel host does) must stay on the same shift stream."""
  hp = _hp(m=10, signal_shape=(1024, 20), n_critic=2)
  B, nd, nc = 4, hp.noise_dim, hp.n_critic
  real, noises, alphas, _ = O.synthetic_batch(hp, B, seed=1, n_critic=nc)
  r0 = engine(hp, B, mixed=True, world_size=2, rank=0, seed=5)
  r1 = engine(hp, B, mixed=True, world_size=2, rank=1, seed=5)
  for step in range(3):
    r0.train_step(real)                                              # rank 0: whole step, nothing injected
    n0, a0, s0 = r0.last_draws((nc + 1) * B * nd, nc * B, 12 * nc + 4)
    s1, n1 = [], []
    for i in range(nc):                                              # rank 1: sub-steps, injects noise on odd steps
      inj = noises[i] if step % 2 else None
      r1.critic_step(real, inj, None, None, update=False, sync=False)
      n, a, s = r1.last_draws(B * nd, B, 12)
      s1.append(s.copy()); n1.append(n.cpu().numpy())
      r1.apply_update(L.DISCRIMINATOR)
    r1.generator_step(real, None, None, update=False, sync=False)
    n, _, s = r1.last_draws(B * nd, 0, 4)
    s1.append(s.copy()); n1.append(n.cpu().numpy())
    r1.apply_update(L.GENERATOR)
    np.testing.assert_array_equal(np.concatenate(s1), s0)
    if step % 2 == 0:
      assert not np.array_equal(np.concatenate(n1), n0.cpu().numpy())
      assert abs(np.corrcoef(np.concatenate(n1), n0.cpu().numpy())[0, 1]) < 0.2
  # same rank, same seed, same calls -> same noise (streams are functions of the call sequence only)
  r1b = engine(hp, B, mixed=True, world_size=2, rank=1, seed=5)
  r1b.critic_step(real, None, None, None, update=False)
  r1c = engine(hp, B, mixed=True, world_size=2, rank=1, seed=5)
  r1c.critic_step(real, None, None, None, update=False)
  assert torch.equal(r1b.last_draws(B * nd, B, 0)[0], r1c.last_draws(B * nd, B, 0)[0])
