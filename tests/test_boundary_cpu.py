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
  assert C.sizeof(_lib.CgConfig) == 4 * 12 + 4 * 4 + 4 * 3 + 4 * 7


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


def test_registries_keep_reference_behaviour(capsys):
  from calciumgan_b200.algorithms.registry import get_algorithm
  from calciumgan_b200.models.registry import get_models
  with pytest.raises(SystemExit):
    get_models(argparse.Namespace(model='wavegan'), None)     # main.py:242 default is unregistered
  assert 'models wavegan not found' in capsys.readouterr().out
  with pytest.raises(SystemExit):
    get_algorithm(argparse.Namespace(algorithm='lswgan'), None, None, None)
  assert 'Algorithm lswgan not found' in capsys.readouterr().out
