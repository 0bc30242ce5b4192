"""Is the x_hat group's X[l] (GP linearised forward) reproducible between two IDENTICAL engines? Its input v_0 is scaled by
1/||g|| whose sum of squares is accumulated with fp32 atomics. Compares fused vs a second fused engine and fused vs unfused,
with the CTA-pair weight-gradient kernel on and off (CG_WG_NO_PAIR is read at every launch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import calciumgan_oracle as O
from calciumgan_b200 import _lib as L
from tests.test_phase_shuffle_gpu import build, _group_shifts

B = 5
hp = O.HParams()
real, noises, alphas, _ = O.synthetic_batch(hp, B, seed=B, n_critic=1)
fused = build(hp, B)
fused2 = build(hp, B)
unfused = build(hp, B, debug_flags=L.DEBUG_NO_PS_FUSE)
for e in (fused2, unfused):
  e.generator.set_weights(fused.generator.get_weights())
  e.discriminator.set_weights(fused.discriminator.get_weights())
for mode in ('pair', 'nopair', 'pair', 'nopair'):
  if mode == 'nopair': os.environ['CG_WG_NO_PAIR'] = '1'
  else: os.environ.pop('CG_WG_NO_PAIR', None)
  bad = {'fused2': 0, 'unfused': 0}
  worst = 0.0
  n = 0
  for rep in range(6):
    for s in (-10, -1, 4, 10):
      sh = _group_shifts(s).reshape(-1)
      for e in (fused, fused2, unfused):
        e.engine.critic_step(real, noises[0], alphas[0], sh, update=False)
      n += 1
      for l in range(1, 6):
        a = fused.engine.debug_read(L.BUF_X, l, 3 * B).float()
        for name, e in (('fused2', fused2), ('unfused', unfused)):
          b = e.engine.debug_read(L.BUF_X, l, 3 * B).float()
          if not torch.equal(a, b):
            bad[name] += 1
            d = (a - b).abs()
            worst = max(worst, float((d / (b.abs() + 1e-30)).max()))
            if bad[name] <= 2:
              print('  %s %s s=%d l=%d: %d of %d elements differ, groups %s' % (mode, name, s, l, int((d > 0).sum()), d.numel(),
                    [int((d[g * B:(g + 1) * B] > 0).sum()) for g in range(3)]))
  print('%s: %d critic steps; layers differing fused-vs-fused2 %d, fused-vs-unfused %d; worst rel diff %.2e' % (mode, n, bad['fused2'], bad['unfused'], worst))
