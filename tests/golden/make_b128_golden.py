"""Generates tests/golden/paper_b128.npz: BASELINE.json configs[1] EXACTLY as benchmarked (paper architecture, batch 128,
2048 x 102) evaluated by the float64 CPU oracle (free-running LeakyReLU branches): scalars, critic scores, the L2 norm
of every per-parameter gradient and a strided sample of every gradient tensor, for one critic step and one generator
step on seeded weights / inputs / draws.  ~4 minutes and ~25 GB of host memory on 8 cores; run once here, the fixture
travels to the GPU box (tests/test_parity_gpu.py::test_headline_batch_128_against_cpu_fixture).

    python tests/golden/make_b128_golden.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import calciumgan_oracle as O   # noqa: E402

BATCH, SEED, STRIDE = 128, 3, 997
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'paper_b128.npz')


def inputs():
  hp = O.HParams()
  gw, dw = O.init_weights(hp, seed=SEED)
  gw, dw = O.randomize_weights(gw, SEED + 1), O.randomize_weights(dw, SEED + 2)
  real, noises, alphas, shifts = O.synthetic_batch(hp, BATCH, seed=SEED + 3, n_critic=1)
  return hp, gw, dw, real, noises, alphas, shifts


def sample_stride(n):
  return STRIDE if n > 8192 else 1     # bias-sized tensors are kept whole


def sample(t):
  return t.reshape(-1)[::sample_stride(t.numel())].numpy().astype(np.float64)


def main():
  torch.set_num_threads(os.cpu_count() or 1)
  hp, gw, dw, real, noises, alphas, shifts = inputs()
  t0 = time.time()
  c = O.critic_step(gw, dw, real, noises[0], alphas[0], shifts[:12].reshape(3, 4), hp)
  print('critic step: %.0f s' % (time.time() - t0))
  t0 = time.time()
  g = O.generator_step(gw, dw, real, noises[1], shifts[12:16], hp)
  print('generator step: %.0f s' % (time.time() - t0))
  out = {
      'c_scalars': np.array([c['dis_loss'], c['gradient_penalty']]),
      'c_scores': torch.cat([c['real_out'], c['fake_out']]).reshape(-1).numpy(),
      'c_gp_norm': c['gp_norm'].numpy(),
      'c_fake_sample': c['fake'].reshape(-1)[::100003].numpy(),
      'c_grad_norms': np.array([float(x.norm()) for x in c['grads']]),
      'g_scalars': np.array([g['gen_loss']] + [g['metrics'][k] for k in sorted(g['metrics'])]),
      'g_scores': g['fake_out'].reshape(-1).numpy(),
      'g_grad_norms': np.array([float(x.norm()) for x in g['grads']]),
  }
  for i, x in enumerate(c['grads']):
    out['c_grad%02d' % i] = sample(x)
  for i, x in enumerate(g['grads']):
    out['g_grad%02d' % i] = sample(x)
  np.savez_compressed(OUT, **out)
  print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
  main()
