"""Generate tests/golden/reference_step.npz by running the reference's own code (files under /root/reference, unmodified)
over the torch-backed TensorFlow stand-in (oracle/tf_shim, float64): one WGAN_GP.train step (n_critic = 2), one
validate and one generate on seeded weights / inputs / draws.

  python tests/golden/make_reference_golden.py        # needs /root/reference (build container only)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import calciumgan_oracle as O  # noqa: E402  (HParams + seeded inputs only)
from oracle import reference_runner as R  # noqa: E402

HP = dict(signal_shape=(128, 12), noise_dim=4, num_units=8, kernel_size=24, m=3, n_critic=2)
BATCH, SEED = 3, 1


def inputs():
  hp = O.HParams(**HP)
  gw, dw = O.init_weights(hp, seed=SEED)
  gw, dw = O.randomize_weights(gw, SEED + 1), O.randomize_weights(dw, SEED + 2)
  real, noises, alphas, shifts = O.synthetic_batch(hp, BATCH, seed=SEED + 3)
  return hp, gw, dw, real, noises, alphas, shifts


# a second, larger case (critic layers with >= 128 time rows: the slab-mode tensor-core kernels, fused PhaseShuffle in
# both directions, m = 10): outputs only, no weights, to keep the fixture small
HP_MEDIUM = dict(signal_shape=(512, 102), noise_dim=8, num_units=32, kernel_size=24, m=10, n_critic=1)
BATCH_MEDIUM, SEED_MEDIUM = 4, 7


def inputs_medium():
  hp = O.HParams(**HP_MEDIUM)
  gw, dw = O.init_weights(hp, seed=SEED_MEDIUM)
  gw, dw = O.randomize_weights(gw, SEED_MEDIUM + 1), O.randomize_weights(dw, SEED_MEDIUM + 2)
  real, noises, alphas, shifts = O.synthetic_batch(hp, BATCH_MEDIUM, seed=SEED_MEDIUM + 3, n_critic=1)
  shifts = np.asarray(shifts).copy()
  shifts[:12] = [10, -10, 7, -3, -10, 10, -1, 4, 9, -9, 10, -10]
  return hp, gw, dw, real, noises, alphas, shifts


def grad_stride(n):
  return 97 if n > 4096 else 1


def main_medium():
  hp, gw, dw, real, noises, alphas, shifts = inputs_medium()
  out = {}
  mods, gan = R.build(hp, BATCH_MEDIUM, gw, dw)
  mods.tf.random.inject(normal=[noises[0]], uniform=[alphas[0]], ints=[int(s) for s in shifts[:12]])
  fake, gen_loss, dis_loss, gp, metrics = gan.validate(torch.as_tensor(real, dtype=torch.float64))
  out['val_fake'] = fake.detach().numpy().astype(np.float32)
  out['val_scalars'] = np.array([float(gen_loss.detach()), float(dis_loss.detach()), float(gp.detach())] +
                                [float(metrics[k].detach()) for k in sorted(metrics)])
  r = R.train_step(hp, gw, dw, real, noises, alphas, shifts)
  out['train_scalars'] = np.array([r['gen_loss'], r['dis_loss'], r['gradient_penalty']] +
                                  [r['metrics'][k] for k in sorted(r['metrics'])])
  # per-tensor norms of the weight updates: a compact check of every gradient's scale after Adam
  out['gen_update_norms'] = np.array([np.linalg.norm(a - b) for a, b in zip(r['gen_weights'], gw)])
  out['dis_update_norms'] = np.array([np.linalg.norm(a - b) for a, b in zip(r['dis_weights'], dw)])
  # per-parameter gradients of the reference's own sub-steps on the initial weights: norms + a strided sample
  sg = R.sub_step_gradients(hp, gw, dw, real, noises[0], alphas[0], shifts[:12], noises[1], shifts[12:16])
  out['sub_scalars'] = np.array([sg['dis_loss'], sg['gradient_penalty'], sg['gen_loss']])
  for key, grads in (('c', sg['dis_grads']), ('g', sg['gen_grads'])):
    out['%s_grad_norms' % key] = np.array([np.linalg.norm(x) for x in grads])
    for i, x in enumerate(grads):
      out['%s_grad_%02d' % (key, i)] = x.reshape(-1)[::grad_stride(x.size)].astype(np.float64)
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_medium.npz')
  np.savez_compressed(path, **out)
  print('wrote', path)


def main():
  main_medium()
  hp, gw, dw, real, noises, alphas, shifts = inputs()
  out = {}
  # validate + generate on the initial weights (gan.py:87-97)
  mods, gan = R.build(hp, BATCH, gw, dw)
  mods.tf.random.inject(normal=[noises[0]], uniform=[alphas[0]], ints=[int(s) for s in shifts[:12]])
  fake, gen_loss, dis_loss, gp, metrics = gan.validate(torch.as_tensor(real, dtype=torch.float64))
  out['val_fake'] = fake.detach().numpy()
  out['val_scalars'] = np.array([float(gen_loss.detach()), float(dis_loss.detach()), float(gp.detach())] +
                                [float(metrics[k].detach()) for k in sorted(metrics)])
  out['gen_fake'] = gan.generate(torch.as_tensor(noises[1], dtype=torch.float64)).detach().numpy()
  out['gen_fake_denorm'] = gan.generate(torch.as_tensor(noises[1], dtype=torch.float64), denorm=True).detach().numpy()
  # one full train step (wgan_gp.py:82-95)
  r = R.train_step(hp, gw, dw, real, noises, alphas, shifts)
  out['train_scalars'] = np.array([r['gen_loss'], r['dis_loss'], r['gradient_penalty']] +
                                  [r['metrics'][k] for k in sorted(r['metrics'])])
  for i, w in enumerate(r['gen_weights']):
    out['gen_w_%02d' % i] = w
  for i, w in enumerate(r['dis_weights']):
    out['dis_w_%02d' % i] = w
  out['draw_order'] = np.array(r['draw_order'])
  # per-parameter gradients (optimizer.py:31-34) of the reference's own critic / generator sub-steps on the initial weights
  sg = R.sub_step_gradients(hp, gw, dw, real, noises[0], alphas[0], shifts[:12], noises[1], shifts[12:16])
  out['sub_scalars'] = np.array([sg['dis_loss'], sg['gradient_penalty'], sg['gen_loss']])
  for i, x in enumerate(sg['dis_grads']):
    out['c_grad_%02d' % i] = x.astype(np.float32)
  for i, x in enumerate(sg['gen_grads']):
    out['g_grad_%02d' % i] = x.astype(np.float32)
  # ... and of the LAST critic update / the generator update inside the full train step above
  for i, x in enumerate(r['dis_grads'][-1]):
    out['t_c_grad_%02d' % i] = x.astype(np.float32)
  for i, x in enumerate(r['gen_grads'][-1]):
    out['t_g_grad_%02d' % i] = x.astype(np.float32)
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_step.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, {k: v.shape for k, v in out.items() if k.endswith('scalars')})


if __name__ == '__main__':
  main()
