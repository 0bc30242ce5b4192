"""PhaseShuffle on the hot path is bit-exact (north_star: "Phase-shuffle indexing must be bit-exact").

The bf16 tensor-core path never runs the stand-alone gather kernel: the forward gather is written in scatter form from
the conv GEMM epilogue (cg_kernels_tc.cuh, ps_scatter_targets) and its adjoint is folded into the data-gradient
epilogue (EPI_PS_MASK, ps_adjoint_row). These tests check THOSE code paths:
  * forward: X[l] (what the next conv reads) == H[l] (what the conv wrote) gathered along the reference's
    pad-and-slice map (gan/models/calciumgan.py:117-138), element for element, for every shift in [-10, 10], every
    layer width of the paper architecture (1024, 512, 256, 128), batch 1 / 3 / 128 and three call groups with
    different shifts; and == the unfused engine (CG_DEBUG_NO_PS_FUSE).
  * GP linearised forward (EPI_MASK + scatter): fused == unfused, element for element.
  * adjoint: with small-integer gradients and weights every sum is exact in fp32 and representable in bf16, so
    fused == unfused == float64 scatter-add, bit for bit.
The index arithmetic itself is additionally checked on the host for every (w, shift) in tests/test_boundary_cpu.py."""
import numpy as np
import pytest
import torch

from oracle import calciumgan_oracle as O
from calciumgan_b200 import _lib as L
from tests.util import namespace_from_oracle

pytestmark = pytest.mark.gpu


def build(hp, batch, mixed=True, **kw):
  from calciumgan_b200.algorithms.registry import get_algorithm
  from calciumgan_b200.models.registry import get_models
  ns = namespace_from_oracle(hp, batch, mixed_precision=mixed, **kw)
  g, d = get_models(ns, None)
  return get_algorithm(ns, g, d, None)


def _gather(h, w, shift):
  idx = torch.from_numpy(O.phase_shuffle_index(w, int(shift)).astype(np.int64)).to(h.device)
  return h.index_select(1, idx)


def _group_shifts(s):
  """12 shifts (3 critic calls x 4 layers) that put different values of [-10, 10] on the three groups"""
  a = [s, -s, ((s + 17) % 21) - 10]
  return np.array([[a[g], a[(g + 1) % 3], a[(g + 2) % 3], a[g]] for g in range(3)], np.int32)


@pytest.mark.parametrize('B,shift_values', [(1, range(-10, 11)), (3, range(-10, 11)), (128, (-10, -3, 0, 1, 10))])
def test_fused_forward_gather_is_bit_exact(B, shift_values):
  hp = O.HParams()
  rng = np.random.RandomState(B)
  real = rng.uniform(0, 1, size=(B, 2048, 102)).astype(np.float32)
  noise = rng.standard_normal((B, hp.noise_dim)).astype(np.float32)
  alpha = rng.uniform(0, 1, size=(B,)).astype(np.float32)
  fused = build(hp, B)
  unfused = build(hp, B, debug_flags=L.DEBUG_NO_PS_FUSE)
  unfused.generator.set_weights(fused.generator.get_weights())
  unfused.discriminator.set_weights(fused.discriminator.get_weights())
  n_fused = n_unfused = 0
  for s in shift_values:
    sh = _group_shifts(s)
    l0 = fused.engine.launch_count()
    fused.validate(real, noise=noise, alpha=alpha, shifts=sh.reshape(-1))
    n_fused = fused.engine.launch_count() - l0
    l0 = unfused.engine.launch_count()
    unfused.validate(real, noise=noise, alpha=alpha, shifts=sh.reshape(-1))
    n_unfused = unfused.engine.launch_count() - l0
    for l in range(1, 5):
      w = 2048 >> l
      x = fused.engine.debug_read(L.BUF_X, l, 3 * B)
      h = fused.engine.debug_read(L.BUF_H, l, 3 * B)
      for g in range(3):
        want = _gather(h[g * B:(g + 1) * B], w, sh[g, l - 1])
        assert torch.equal(x[g * B:(g + 1) * B], want), (B, s, l, g)
        if B <= 3:    # the literal pad-and-slice transcription of calciumgan.py:126-137
          np.testing.assert_array_equal(x[g * B:(g + 1) * B].cpu().numpy(),
                                        O.phase_shuffle_literal(h[g * B:(g + 1) * B].cpu().numpy(), int(sh[g, l - 1])))
      assert torch.equal(x, unfused.engine.debug_read(L.BUF_X, l, 3 * B)), (B, s, l)
      assert torch.equal(h, unfused.engine.debug_read(L.BUF_H, l, 3 * B)), (B, s, l)
  assert n_unfused == n_fused + 4    # the four stand-alone gather launches exist only in the unfused engine
  assert fused.engine.tc_launch_count() > 0


@pytest.mark.parametrize('B', [2, 5])
def test_fused_gather_of_the_linearised_gp_forward_is_bit_exact(B):
  """critic step: v_l = PS(M_l * conv(v_{l-1})) lands in the x_hat group's X[l] slots (wgan_gp.py:43-50 via the 4-pass
  gradient penalty): fused epilogue == conv + stand-alone gather."""
  hp = O.HParams()
  real, noises, alphas, _ = O.synthetic_batch(hp, B, seed=B, n_critic=1)
  fused = build(hp, B)
  unfused = build(hp, B, debug_flags=L.DEBUG_NO_PS_FUSE)
  unfused.generator.set_weights(fused.generator.get_weights())
  unfused.discriminator.set_weights(fused.discriminator.get_weights())
  for s in (-10, -1, 4, 10):
    sh = _group_shifts(s).reshape(-1)
    fused.engine.critic_step(real, noises[0], alphas[0], sh, update=False)
    unfused.engine.critic_step(real, noises[0], alphas[0], sh, update=False)
    for l in range(1, 6):
      assert torch.equal(fused.engine.debug_read(L.BUF_X, l, 3 * B), unfused.engine.debug_read(L.BUF_X, l, 3 * B)), (s, l)
    for a, b in zip(fused.engine.get_grads(L.DISCRIMINATOR), unfused.engine.get_grads(L.DISCRIMINATOR)):
      # same kernels downstream of identical inputs; the weight-gradient kernels accumulate their row splits with
      # fp32 atomics, so two runs agree to summation order, not bit for bit
      assert np.abs(a - b).max() <= 1e-5 * max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize('layer', [2, 3, 4])
def test_fused_adjoint_is_bit_exact_on_integers(layer):
  """DA[l] -> DA[l-1] (data gradient, PhaseShuffle adjoint, LeakyReLU slope) with integer-valued gradients and weights:
  every partial sum is an integer below 2^8, exact in fp32 and in bf16, so the fused epilogue (reflected rows summed
  in registers), the two-kernel form and float64 scatter-add along the reference's index map agree bit for bit."""
  hp = O.HParams()
  B, group_b = 6, 2
  fused = build(hp, B)
  unfused = build(hp, B, debug_flags=L.DEBUG_NO_PS_BWD_FUSE)
  rng = np.random.RandomState(layer)
  dw = fused.discriminator.get_weights()
  for i in range(0, 10, 2):   # sparse +-1 kernels: a 12-tap x C_out sum stays far below 256
    dw[i] = (rng.randint(-1, 2, size=dw[i].shape) * (rng.uniform(size=dw[i].shape) < 1.0 / 32)).astype(np.float32)
  fused.discriminator.set_weights(dw)
  unfused.discriminator.set_weights(dw)
  dc = O.discriminator_channels(hp)
  w_out, w_in = 2048 >> layer, 2048 >> (layer - 1)
  dy = (rng.randint(-2, 3, size=(B, w_out, dc[layer])) * (rng.uniform(size=(B, w_out, dc[layer])) < 0.25)).astype(np.float32)
  h = rng.standard_normal((B, w_in, dc[layer - 1])).astype(np.float32)
  kernel = torch.tensor(dw[2 * (layer - 1)], dtype=torch.float64)
  for s in range(-10, 11):
    shifts = [s, -s, ((s + 17) % 21) - 10]
    a = fused.engine.dgrad_ps(layer, dy, h, group_b, shifts)
    b = unfused.engine.dgrad_ps(layer, dy, h, group_b, shifts)
    assert torch.equal(a, b), (layer, s)
    dx = O.conv1d_same_dgrad(torch.tensor(dy, dtype=torch.float64), kernel, w_in)
    acc = torch.cat([O._ps_transpose(dx[g * group_b:(g + 1) * group_b], shifts[g]) for g in range(3)])
    assert float(acc.abs().max()) < 256 and float(acc.abs().max()) > 0
    slope = torch.where(torch.tensor(h).bfloat16().float() > 0, torch.tensor(1.0), torch.tensor(O.LEAKY_ALPHA, dtype=torch.float32))
    want = (acc.float() * slope).bfloat16().float()   # fp32 product, one rounding to bf16: the engine's storage
    assert torch.equal(a.cpu(), want), (layer, s, float((a.cpu() - want).abs().max()))
  assert fused.engine.tc_launch_count() > 0


def test_gradient_penalty_activations_do_not_depend_on_atomic_order():
  """v_0 = u(||g||) * g feeds the linearised pass, and ||g||^2 is accumulated with atomics from many CTAs. In fp32 its last
  bit followed their arrival order, so two identical engines occasionally disagreed in thousands of activations
  (profiles/r2_gp_reproducibility.txt); the double accumulator adds the fp32 partials exactly. Two identical engines must
  now agree on every X[l] of every critic step, element for element."""
  hp = O.HParams()
  B = 5
  real, noises, alphas, _ = O.synthetic_batch(hp, B, seed=B, n_critic=1)
  a = build(hp, B)
  b = build(hp, B)
  b.generator.set_weights(a.generator.get_weights())
  b.discriminator.set_weights(a.discriminator.get_weights())
  for rep in range(3):
    for s in (-10, -1, 4, 10):
      sh = _group_shifts(s).reshape(-1)
      a.engine.critic_step(real, noises[0], alphas[0], sh, update=False)
      b.engine.critic_step(real, noises[0], alphas[0], sh, update=False)
      for l in range(1, 6):
        assert torch.equal(a.engine.debug_read(L.BUF_X, l, 3 * B), b.engine.debug_read(L.BUF_X, l, 3 * B)), (rep, s, l)
