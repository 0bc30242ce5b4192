"""Kernel-level parity: every conv layer's forward / data-gradient / weight-gradient GEMM in
isolation through cg_debug_layer, against the fp64 oracle convolutions.  No LeakyReLU slope is
involved in dgrad/wgrad and inputs are bf16-representable, so the bf16 tensor-core kernels can be
held to the rounding of their own output (bf16: 2^-9 per element; fp32 outputs: accumulation only)."""
import numpy as np
import pytest
import torch

from oracle import calciumgan_oracle as O
from tests.util import namespace_from_oracle, rel_err

pytestmark = pytest.mark.gpu

CONFIGS = {
    'medium': (dict(signal_shape=(512, 102), noise_dim=8, num_units=32, kernel_size=24, m=3, n_critic=1), 6),
    'tiny_ragged': (dict(signal_shape=(64, 6), noise_dim=4, num_units=4, kernel_size=6, m=1, n_critic=1), 5),
    'wide_odd_k': (dict(signal_shape=(256, 70), noise_dim=16, num_units=64, kernel_size=5, m=2, n_critic=1), 3),
    # BASELINE.json configs[1] architecture (row-pair form of conv1, slab / per-tap modes of every other layer)
    'paper': (dict(signal_shape=(2048, 102), noise_dim=32, num_units=64, kernel_size=24, m=10, n_critic=1), 2),
    # BASELINE.json configs[3] widths (num_units 128, 512 channels: N = 640 / 512 tiles, K chunks up to 640, layer norm
    # over 640 channels outside the GEMM epilogue) at a shorter sequence
    'scaled_widths': (dict(signal_shape=(1024, 512), noise_dim=32, num_units=128, kernel_size=24, m=10, n_critic=1), 2),
}


def _bf(x):
  return torch.as_tensor(x, dtype=torch.float32).to(torch.bfloat16).to(torch.float64)


def _engine(cfg, mixed, force_simt=False):
  from calciumgan_b200.models.registry import get_models
  kw, B = CONFIGS[cfg]
  if cfg == 'scaled_widths' and not mixed:
    pytest.skip('the fp32 debug path keeps a layer-norm row in registers: at most 512 channels (config 4 is a bf16 configuration)')
  hp = O.HParams(**kw)
  ns = namespace_from_oracle(hp, B, mixed_precision=mixed, force_simt=force_simt)
  g, d = get_models(ns, None)
  gw, dw = O.init_weights(hp, seed=5)
  gw, dw = O.randomize_weights(gw, 6), O.randomize_weights(dw, 7)
  # bf16-representable kernels so weight rounding is not part of the comparison
  gw = [_bf(a).float().numpy() if a.ndim > 1 else a for a in gw]
  dw = [_bf(a).float().numpy() if a.ndim > 1 else a for a in dw]
  g.set_weights(gw)
  d.set_weights(dw)
  return hp, B, g.engine, gw, dw


def _tols(mixed):
  # bf16 outputs carry one rounding (<= 2^-9 relative per element); fp32 outputs only accumulation error
  return (4e-3, 2e-4) if mixed else (1e-5, 1e-5)


@pytest.mark.parametrize('cfg', list(CONFIGS))
@pytest.mark.parametrize('mode', ['fp32', 'bf16_simt', 'bf16_tc'])
def test_critic_layers(cfg, mode):
  mixed = mode != 'fp32'
  hp, B, eng, gw, dw = _engine(cfg, mixed, force_simt=(mode == 'bf16_simt'))
  tol_act, tol_w = _tols(mixed)
  rng = np.random.RandomState(0)
  dc = O.discriminator_channels(hp)
  L, K = hp.signal_shape[0], hp.kernel_size
  for l in range(1, 6):
    lin, lout, ci, co = L >> (l - 1), L >> l, dc[l - 1], dc[l]
    for Bt in sorted({B, 3 * B}):
      x = _bf(rng.standard_normal((Bt, lin, ci)))
      dy = _bf(rng.standard_normal((Bt, lout, co)))
      w = torch.tensor(dw[2 * (l - 1)], dtype=torch.float64)
      b = torch.tensor(dw[2 * (l - 1) + 1], dtype=torch.float64)
      ref = O.leaky_relu(O.conv1d_same(x, w, b))
      got = eng.debug_layer(1, l, 0, x=x.float()).cpu().numpy()
      assert rel_err(got, ref.numpy()) <= tol_act, ('fwd', cfg, mode, l, Bt, rel_err(got, ref.numpy()))
      if l == 1 and Bt > B:
        pass   # the layer-1 data gradient buffer holds max_batch samples only
      else:
        ref = O.conv1d_same_dgrad(dy, w, lin)
        got = eng.debug_layer(1, l, 1, dy=dy.float()).cpu().numpy()
        assert rel_err(got, ref.numpy()) <= tol_act, ('dgrad', cfg, mode, l, Bt, rel_err(got, ref.numpy()))
      ref = O.conv1d_same_wgrad(x, dy, K)
      got = eng.debug_layer(1, l, 2, x=x.float(), dy=dy.float()).cpu().numpy()
      assert rel_err(got, ref.numpy()) <= tol_w, ('wgrad', cfg, mode, l, Bt, rel_err(got, ref.numpy()))
  # the tensor-core kernels must be what ran in bf16_tc mode (and only there)
  assert (eng.tc_launch_count() > 0) == (mode == 'bf16_tc'), (mode, eng.tc_launch_count())


@pytest.mark.parametrize('cfg', list(CONFIGS))
@pytest.mark.parametrize('mode', ['fp32', 'bf16_simt', 'bf16_tc'])
def test_generator_layers(cfg, mode):
  mixed = mode != 'fp32'
  hp, B, eng, gw, dw = _engine(cfg, mixed, force_simt=(mode == 'bf16_simt'))
  tol_act, tol_w = _tols(mixed)
  rng = np.random.RandomState(1)
  gc = O.generator_channels(hp)
  w0, K = hp.signal_shape[0] // 32, hp.kernel_size
  for i in range(1, 6):
    lin, lout, ci, co = w0 << (i - 1), w0 << i, gc[i - 1], gc[i]
    x = _bf(rng.standard_normal((B, lin, ci)))
    dy = _bf(rng.standard_normal((B, lout, co)))
    idx = 2 + (i - 1) * 4
    w = torch.tensor(gw[idx], dtype=torch.float64)          # (K, 1, Cout, Cin)
    b = torch.tensor(gw[idx + 1], dtype=torch.float64)
    ref = O.conv1d_transpose_same(x, w, b)
    got = eng.debug_layer(0, i, 0, x=x.float()).cpu().numpy()
    assert rel_err(got, ref.numpy()) <= tol_act, ('fwd', cfg, mode, i, rel_err(got, ref.numpy()))
    ref = O.conv1d_same(dy, w[:, 0], None)                   # data gradient of the transposed conv
    got = eng.debug_layer(0, i, 1, dy=dy.float()).cpu().numpy()
    assert rel_err(got, ref.numpy()) <= tol_act, ('dgrad', cfg, mode, i, rel_err(got, ref.numpy()))
    ref = O.conv1d_same_wgrad(dy, x, K).reshape(K, 1, co, ci)
    got = eng.debug_layer(0, i, 2, x=x.float(), dy=dy.float()).cpu().numpy()
    assert rel_err(got, ref.numpy()) <= tol_w, ('wgrad', cfg, mode, i, rel_err(got, ref.numpy()))
