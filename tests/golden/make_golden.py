"""Generate tests/golden/tiny_step.npz from the CPU oracle (fp64).

The reference ships no golden vectors and cannot be imported here (TensorFlow 2.3.1 absent),
so these fixtures are ORACLE-generated: they pin the oracle against regressions and give the
GPU parity tests a fixed target, but they do not pin the oracle to TensorFlow ("parity
unpinned", see oracle/calciumgan_oracle.py).

  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

from oracle import calciumgan_oracle as O  # noqa: E402

BATCH = 3
N_CRITIC = 2


def tiny_hp():
  return O.HParams(signal_shape=(64, 6), noise_dim=4, num_units=4, kernel_size=6, m=2,
                   n_critic=N_CRITIC)


def compute():
  hp = tiny_hp()
  gw, dw = O.init_weights(hp, seed=1234)
  gw, dw = O.randomize_weights(gw, 1), O.randomize_weights(dw, 2)
  real, noises, alphas, shifts = O.synthetic_batch(hp, BATCH, seed=1234)
  out = {'real': real, 'noises': noises, 'alphas': alphas, 'shifts': shifts}
  for i, a in enumerate(gw):
    out['gen_w%02d' % i] = a
  for i, a in enumerate(dw):
    out['dis_w%02d' % i] = a
  c = O.critic_step(gw, dw, real, noises[0], alphas[0], shifts[:12].reshape(3, 4), hp)
  out['c_fake'] = c['fake'].numpy()
  out['c_real_out'] = c['real_out'].numpy()
  out['c_fake_out'] = c['fake_out'].numpy()
  out['c_gp_grad'] = c['gp_grad'].numpy()
  out['c_scalars'] = np.array([c['dis_loss'], c['gradient_penalty']])
  for i, g in enumerate(c['grads']):
    out['c_grad%02d' % i] = g.numpy()
  g = O.generator_step(gw, dw, real, noises[N_CRITIC], shifts[12 * N_CRITIC:12 * N_CRITIC + 4], hp)
  out['g_scalars'] = np.array([g['gen_loss']] + [g['metrics'][k] for k in sorted(g['metrics'])])
  for i, x in enumerate(g['grads']):
    out['g_grad%02d' % i] = x.numpy()
  st = O.TrainState.create(gw, dw)
  gen_loss, dis_loss, gp, metrics = O.train_step(st, real, noises, alphas, shifts, hp)
  out['t_scalars'] = np.array([gen_loss, dis_loss, gp] + [metrics[k] for k in sorted(metrics)])
  for i, a in enumerate(st.gen):
    out['t_gen_w%02d' % i] = a.numpy()
  for i, a in enumerate(st.dis):
    out['t_dis_w%02d' % i] = a.numpy()
  return out


if __name__ == '__main__':
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tiny_step.npz')
  np.savez_compressed(path, **compute())
  print('wrote', path, os.path.getsize(path), 'bytes')
