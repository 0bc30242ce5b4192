"""Data-parallel host logic on CPU: world_size-2 gloo processes drive WGAN_GP._train_dp over a stub
engine (the CUDA engine has no CPU path by design). Checks the contract of SURVEY §8e: gradients are
summed across ranks before the optimizer, every rank applies the same update, scalars are averaged,
and all ranks receive the same PhaseShuffle draws."""
import argparse
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class StubEngine(object):
  """Mimics calciumgan_b200.engine.Engine's surface used by WGAN_GP (CPU tensors)."""

  def __init__(self, rank, n_critic):
    self.rank = rank
    self.device = torch.device('cpu')
    self.cfg = argparse.Namespace(n_critic=n_critic, phase_m=2, noise_dim=4)
    self.g = {0: torch.zeros(8), 1: torch.zeros(6)}
    self.scal = torch.zeros(16)
    self.applied, self.seen_shifts, self.calls = [], [], []

  def to_device(self, x, shape=None):
    return None if x is None else torch.as_tensor(np.asarray(x), dtype=torch.float32)

  def grad_tensor(self, which):
    return self.g[which]

  def grad_buckets(self, which):
    g = self.g[which]
    return [(g[4:], 0), (g[2:4], 1), (g[:2], 2)]     # last layers first, contiguous views of the flat buffer

  def num_buckets(self, which):
    return 3

  def scalars_tensor(self):
    return self.scal

  def prefetch_generator(self, real, noise=None, alpha=None, for_generator_step=False, want_fake32=True):
    self.calls.append('prefetch_g' if for_generator_step else 'prefetch_c')

  def critic_step(self, real, noise, alpha, shifts, update=True, sync=True, same_real=False, want_fake32=True,
                  gen_prefetched=False):
    assert not update and not sync and not want_fake32
    self.calls.append('critic+' if gen_prefetched else 'critic')
    self.seen_shifts.append(None if shifts is None else np.asarray(shifts).copy())
    self.g[1][:] = float(real.sum()) * torch.arange(1, 7)       # rank-dependent "gradient"
    self.scal[0], self.scal[1] = 10.0 + self.rank, 1.0 + self.rank

  def generator_step(self, real, noise, shifts, update=True, sync=True, gen_prefetched=False):
    assert not update and not sync
    self.calls.append('generator+' if gen_prefetched else 'generator')
    self.seen_shifts.append(None if shifts is None else np.asarray(shifts).copy())
    self.g[0][:] = float(real.mean()) * torch.arange(1, 9)
    self.scal[4] = -3.0 - self.rank
    self.scal[5:9] = torch.tensor([1.0, 2.0, 3.0, 4.0]) * (self.rank + 1)

  def apply_update(self, which):
    self.calls.append('adam%d' % which)
    self.applied.append((which, self.g[which].clone()))


def _worker(rank, world, port, out):
  sys.path.insert(0, ROOT)
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from calciumgan_b200.algorithms.wgan_gp import WGAN_GP
  nc = 2
  eng = StubEngine(rank, nc)
  handle = argparse.Namespace(engine=eng)
  hp = argparse.Namespace(noise_shape=(4,), normalize=True, signals_min=0.0, signals_max=1.0, learning_rate=1e-4,
                          mixed_precision=False, gradient_penalty=10.0, n_critic=nc, conv2d=False)
  gan = WGAN_GP(hp, handle, handle, None)
  real = np.full((2, 4, 3), rank + 1.0, np.float32)            # rank shard
  shifts = np.arange(12 * nc + 4) % 5 - 2
  res = gan.train(real, shifts=shifts)
  out[rank] = dict(res=res, applied=[(w, g.numpy()) for w, g in eng.applied],
                   shifts=[s.tolist() for s in eng.seen_shifts], calls=list(eng.calls))
  dist.destroy_process_group()


def test_dp_two_ranks_gloo():
  world = 2
  mgr = mp.Manager()
  out = mgr.dict()
  port = 29500 + os.getpid() % 2000
  mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
  r0, r1 = out[0], out[1]
  # both ranks applied identical, rank-summed gradients: critic x2 then generator
  assert [w for w, _ in r0['applied']] == [1, 1, 0]
  for (w0, g0), (w1, g1) in zip(r0['applied'], r1['applied']):
    np.testing.assert_array_equal(g0, g1)
  sum_real = 24 * 1.0 + 24 * 2.0
  np.testing.assert_allclose(r0['applied'][0][1], sum_real * np.arange(1, 7))
  np.testing.assert_allclose(r0['applied'][2][1], (1.0 + 2.0) * np.arange(1, 9))
  # scalars are averaged over ranks; dis_loss / gp are means over the critic sub-steps first
  gen_loss, dis_loss, gp, metrics = r0['res']
  assert r1['res'][:3] == r0['res'][:3]
  assert abs(dis_loss - 10.5) < 1e-6 and abs(gp - 1.5) < 1e-6 and abs(gen_loss + 3.5) < 1e-6
  assert abs(metrics['signals_metrics/min'] - 1.5) < 1e-6 and abs(metrics['signals_metrics/std'] - 6.0) < 1e-6
  # identical PhaseShuffle draws on every rank, 12 per critic sub-step then 4
  assert r0['shifts'] == r1['shifts']
  assert [len(s) for s in r0['shifts']] == [12, 12, 4]
  # the generator part of the next sub-step is enqueued before the critic's Adam (it hides the all-reduce tail), and the
  # sub-step that follows is told that its generator forward already ran
  assert r0['calls'] == ['critic', 'prefetch_c', 'adam1', 'critic+', 'prefetch_g', 'adam1', 'generator+', 'adam0'], r0['calls']
