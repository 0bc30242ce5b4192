import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def tiny_hp():
  """Small config every CPU test can afford: L=64, C=6, nu=4, K=6 (w = 2)."""
  from oracle.calciumgan_oracle import HParams
  return HParams(signal_shape=(64, 6), noise_dim=4, num_units=4, kernel_size=6, m=2)
