"""N-rank data-parallel step == 1-rank step on the concatenated batch.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
fp32 mode so the comparison is tight; every rank also holds a single-process engine fed the global batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from oracle import calciumgan_oracle as O
from tests.util import namespace_from_oracle, rel_err
from calciumgan_b200.algorithms.wgan_gp import WGAN_GP
from calciumgan_b200.models.calciumgan import ModelHandle, _discriminator_names, _generator_names
from calciumgan_b200.engine import Engine, hparams_to_config

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
mixed = len(sys.argv) > 1 and sys.argv[1] == 'bf16'
# bf16: 102 channels (padded to 128) so that the tensor-core generator-head kernel and its CG_FLAG_NO_FAKE32 mode are on the path
hp = (O.HParams(signal_shape=(512, 102), noise_dim=8, num_units=32, kernel_size=24, m=3, n_critic=2) if mixed else
      O.HParams(signal_shape=(256, 20), noise_dim=8, num_units=16, kernel_size=24, m=3, n_critic=2))
Bl = 4
ns = namespace_from_oracle(hp, Bl * world, mixed_precision=mixed)


def make(world_size, r):
  eng = Engine(hparams_to_config(ns, world_size=world_size, rank=r))
  g = ModelHandle(eng, 0, 'generator', _generator_names(True))
  d = ModelHandle(eng, 1, 'discriminator', _discriminator_names())
  return eng, g, d


gw, dw = O.init_weights(hp, seed=1)
gw, dw = O.randomize_weights(gw, 2), O.randomize_weights(dw, 3)
real, noises, alphas, shifts = O.synthetic_batch(hp, Bl * world, seed=4)
sl = slice(rank * Bl, (rank + 1) * Bl)

eng, g, d = make(world, rank)
g.set_weights(gw); d.set_weights(dw)
gan = WGAN_GP(ns, g, d, None)
gan.no_dp_overlap = os.environ.get('CG_NO_DP_OVERLAP') is not None
out_dp = gan.train(real[sl], noise=noises[:, sl], alpha=alphas[:, sl], shifts=shifts)

# single-process reference on the concatenated batch (bypass the DP branch)
eng1, g1, d1 = make(1, 0)
g1.set_weights(gw); d1.set_weights(dw)
s = eng1.train_step(real, noises, alphas, shifts)
worst = 0.0
for a, b, w0 in list(zip(d.get_weights(), d1.get_weights(), dw)) + list(zip(g.get_weights(), g1.get_weights(), gw)):
  if np.abs(b - w0).max() > 0:
    if mixed:   # mean absolute deviation of the update relative to its mean size (sign-like Adam steps, see below)
      worst = max(worst, float(np.abs(a - b).mean() / np.abs(b - w0).mean()))
    else:
      worst = max(worst, rel_err(a - w0, b - w0))
# bf16: the first Adam steps are sign-like (lr * g / (|g| + eps)), so near-zero gradient elements whose sign depends on
# the summation order dominate the update error; the losses below are the tight check
tol = 1e-1 if mixed else 2e-3
print('rank %d: dp losses %s | single %s | worst update rel err %.3e' %
      (rank, ['%.5f' % x for x in out_dp[:3]], ['%.5f' % float(s[i]) for i in (4, 0, 1)], worst))
assert worst <= tol, worst
assert abs(out_dp[1] - float(s[0])) <= (5e-2 if mixed else 1e-3) * max(1, abs(float(s[0])))
# a further step with library-drawn randomness: every rank must draw the same PhaseShuffle shifts (and different noise)
gan.train(real[sl])
n_mine, _, sh_mine = eng.last_draws(Bl * hp.noise_dim, 0, 4)
sh_all = [torch.zeros(4, dtype=torch.int32, device='cuda') for _ in range(world)]
dist.all_gather(sh_all, torch.as_tensor(sh_mine, device='cuda'))
assert all(torch.equal(sh_all[0], t) for t in sh_all), 'ranks drew different PhaseShuffle shifts'
n_all = [torch.zeros_like(n_mine) for _ in range(world)]
dist.all_gather(n_all, n_mine)
assert world == 1 or not torch.equal(n_all[0], n_all[1]), 'ranks drew the same noise'
dist.barrier()
if rank == 0:
  print('DP CHECK OK (world %d, %s)' % (world, 'bf16' if mixed else 'fp32'))
dist.destroy_process_group()
