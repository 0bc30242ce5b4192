"""Tensor-core bf16 path vs CUDA-core fp32 path of the same engine on identical inputs, paper architecture, over batch
sizes that exercise ragged tiles / odd CTA pairs / partial waves (no oracle involved: the two paths share no GEMM code)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import calciumgan_oracle as O
from tests.util import namespace_from_oracle
from calciumgan_b200.algorithms.registry import get_algorithm
from calciumgan_b200.models.registry import get_models

hp = O.HParams()
gw, dw = O.init_weights(hp, seed=3)
gw, dw = O.randomize_weights(gw, 4), O.randomize_weights(dw, 5)
worst = 0.0
for B in [int(x) for x in sys.argv[1:]] or [1, 2, 7, 31, 64, 100, 128]:
  real, noises, alphas, shifts = O.synthetic_batch(hp, B, seed=6 + B, n_critic=1)
  res = {}
  for mixed in (False, True):
    ns = namespace_from_oracle(hp, B, mixed_precision=mixed)
    g, d = get_models(ns, None)
    gan = get_algorithm(ns, g, d, None)
    g.set_weights(gw); d.set_weights(dw)
    s = gan.engine.critic_step(real, noises[0], alphas[0], shifts[:12], update=False)
    sc = gan.engine.scores(3 * B).cpu().numpy()[:2 * B]
    fake = gan.engine.fake(B).cpu().numpy()
    s2 = gan.engine.generator_step(real, noises[1], shifts[12:16], update=False)
    res[mixed] = (np.array(s[:4]), sc, fake, np.array(s2[4:9]))
    gan.engine.close()
  a, b = res[True], res[False]
  e = [np.abs(a[0] - b[0]).max() / max(1, np.abs(b[0]).max()), np.abs(a[1] - b[1]).max() / max(1, np.abs(b[1]).max()),
       np.linalg.norm(a[2] - b[2]) / np.linalg.norm(b[2]), np.abs(a[3] - b[3]).max() / max(1, np.abs(b[3]).max())]
  worst = max(worst, max(e))
  print('B=%3d  critic scalars %.1e  scores %.1e  fake %.1e  generator scalars %.1e  %s' %
        (B, e[0], e[1], e[2], e[3], 'ok' if max(e) <= 2e-2 and np.isfinite(max(e)) else 'FAIL'))
assert worst <= 2e-2, worst
print('BATCH SWEEP OK')
