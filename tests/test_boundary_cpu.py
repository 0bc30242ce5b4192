"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/calciumgan_b200.h declares; host logic that needs no GPU; PhaseShuffle index map of
the library is bit-exact against the oracle (int32 index, bit-exact)."""
import argparse
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import calciumgan_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
  src = open(os.path.join(ROOT, 'include', 'calciumgan_b200.h')).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(cg_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_header_symbol():
  from calciumgan_b200 import _lib
  lib = _lib.load()
  names = _header_symbols()
  assert len(names) >= 30
  for n in names:
    assert hasattr(lib, n), 'missing export %s' % n
    assert n in _lib.SIGNATURES, 'ctypes signature missing for %s' % n
  assert lib.cg_version() == 100


def test_config_struct_matches_header_size():
  from calciumgan_b200 import _lib
  assert C.sizeof(_lib.CgConfig) == 4 * 12 + 4 * 4 + 4 * 4 + 4 * 6


def test_create_without_gpu_fails_loudly():
  import torch
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  from calciumgan_b200 import _lib
  from calciumgan_b200.engine import Engine, hparams_to_config
  hp = argparse.Namespace(signal_shape=(64, 6), num_channels=6, noise_dim=4, num_units=4, kernel_size=6,
                          strides=2, m=2, layer_norm=True, normalize=True, batch_size=2, n_critic=1,
                          mixed_precision=False, gradient_penalty=10.0, learning_rate=1e-4)
  with pytest.raises(RuntimeError):
    Engine(hparams_to_config(hp))
  # and straight through the C ABI
  lib = _lib.load()
  ctx = C.c_void_p()
  cfg = hparams_to_config(hp)
  assert lib.cg_create(C.byref(cfg), C.byref(ctx)) != 0
  assert b'no CUDA device' in lib.cg_last_error()


def test_unsupported_hparams_raise():
  from calciumgan_b200.engine import hparams_to_config
  base = dict(signal_shape=(64, 6), num_channels=6, noise_dim=4, num_units=4, kernel_size=6, strides=2,
              m=2, layer_norm=True, normalize=True, batch_size=2)
  with pytest.raises(NotImplementedError):
    hparams_to_config(argparse.Namespace(batch_norm=True, **base))
  with pytest.raises(NotImplementedError):
    hparams_to_config(argparse.Namespace(activation='relu', **base))


def test_noise_shape_error_matches_reference():
  from calciumgan_b200.models.calciumgan import calculate_noise_shape
  assert calculate_noise_shape((2048, 102), 32, 5, 2) == (64, 32)
  with pytest.raises(ValueError, match='is not an integer'):
    calculate_noise_shape((2050, 102), 32, 5, 2)


@pytest.mark.parametrize('w', [4, 64, 128, 1024])
def test_phase_shuffle_index_bit_exact(w):
  from calciumgan_b200.engine import phase_shuffle_index
  m = min(10, w - 1)
  for shift in range(-m, m + 1):
    np.testing.assert_array_equal(phase_shuffle_index(w, shift), O.phase_shuffle_index(w, shift))


def _i32(n):
  a = np.zeros(n, np.int32)
  return a, a.ctypes.data_as(C.POINTER(C.c_int32))


@pytest.mark.parametrize('w', [4, 64, 128, 256, 512, 1024, 2048, 4096])
def test_fused_epilogue_scatter_form_equals_the_gather(w):
  """The tensor-core conv epilogue writes the PhaseShuffle gather in scatter form (ps_scatter_targets, the very
  function the kernel calls): inverted on the host it must be the reference's pad-and-slice map
  (calciumgan.py:117-138) for every shift, every output row written exactly once."""
  from calciumgan_b200 import _lib
  lib = _lib.load()
  m = min(10, w - 1)
  x = np.arange(w, dtype=np.float64).reshape(1, w, 1)
  for shift in range(-m, m + 1):
    (t1, p1), (t2, p2) = _i32(w), _i32(w)
    assert lib.cg_phase_shuffle_scatter_index(w, shift, p1, p2) == 0
    src = np.full(w, -1, np.int64)
    writes = np.zeros(w, np.int64)
    for q in range(w):
      for t in (t1[q], t2[q]):
        if t >= 0:
          src[t] = q
          writes[t] += 1
    assert (writes == 1).all(), (w, shift)
    np.testing.assert_array_equal(src, O.phase_shuffle_index(w, shift))
    np.testing.assert_array_equal(src, O.phase_shuffle_literal(x, shift)[0, :, 0].astype(np.int64))


@pytest.mark.parametrize('w', [128, 256, 512, 1024, 2048])
def test_fused_epilogue_adjoint_plan_equals_scatter_add(w):
  """The data-gradient epilogue's PhaseShuffle adjoint (ps_adjoint_row: store row, exchange slots of rows reflected at
  an edge, rows written as zeros), executed on the host with integer values, equals index_add along the gather map."""
  from calciumgan_b200 import _lib
  lib = _lib.load()
  rng = np.random.RandomState(w)
  for shift in range(-10, 11):
    (dest, pd), (xs, ps), (xp, pp), (zero, pz) = _i32(w), _i32(w), _i32(w), _i32(w)
    assert lib.cg_phase_shuffle_adjoint_plan(w, shift, pd, ps, pp, pz) == 0
    acc = rng.randint(-1000, 1000, size=w).astype(np.int64)
    ref = np.zeros(w, np.int64)
    np.add.at(ref, O.phase_shuffle_index(w, shift).astype(np.int64), acc)
    out = np.full(w, np.iinfo(np.int64).min)
    for par in (0, 1):   # one output phase (row parity) per accumulator block, slots are private to it
      slots = {}
      for t in range(par, w, 2):
        if xs[t] >= 0:
          assert xs[t] not in slots and 0 <= xs[t] < 5 and dest[t] < 0
          slots[xs[t]] = acc[t]
      for t in range(par, w, 2):
        v = acc[t] + (slots[xp[t]] if xp[t] >= 0 else 0)
        if dest[t] >= 0:
          assert dest[t] % 2 == (t + shift) % 2
          assert out[dest[t]] == np.iinfo(np.int64).min
          out[dest[t]] = v
        else:
          assert xs[t] >= 0     # a row that is not stored must be reflected
    for t in range(w):
      if zero[t]:
        assert out[t] == np.iinfo(np.int64).min
        out[t] = 0
    np.testing.assert_array_equal(out, ref)


def test_registries_keep_reference_behaviour(capsys):
  from calciumgan_b200.algorithms.registry import get_algorithm
  from calciumgan_b200.models.registry import get_models
  with pytest.raises(SystemExit):
    get_models(argparse.Namespace(model='wavegan'), None)     # main.py:242 default is unregistered
  assert 'models wavegan not found' in capsys.readouterr().out
  with pytest.raises(SystemExit):
    get_algorithm(argparse.Namespace(algorithm='lswgan'), None, None, None)
  assert 'Algorithm lswgan not found' in capsys.readouterr().out
