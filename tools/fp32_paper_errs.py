"""Per-tensor fp32 errors vs the fp64 oracle (debug aid): python tools/fp32_paper_errs.py B L C nu [layer_norm]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import calciumgan_oracle as O
from tests.util import namespace_from_oracle, rel_err
from calciumgan_b200.algorithms.registry import get_algorithm
from calciumgan_b200.models.registry import get_models
a = [int(x) for x in sys.argv[1:]]
B, L, C, nu = (a + [2, 2048, 102, 64][len(a):])[:4]
ln = a[4] if len(a) > 4 else 1
seed = a[5] if len(a) > 5 else 61
hp = O.HParams(signal_shape=(L, C), num_units=nu, layer_norm=bool(ln))
ns = namespace_from_oracle(hp, B, mixed_precision=False)
g, d = get_models(ns, None); gan = get_algorithm(ns, g, d, None)
gw, dw = O.init_weights(hp, seed=seed)
gw, dw = O.randomize_weights(gw, seed + 1), O.randomize_weights(dw, seed + 2)
g.set_weights(gw); d.set_weights(dw)
real, noises, alphas, shifts = O.synthetic_batch(hp, B, seed=seed + 3, n_critic=1)
ref_g = O.generator_step(gw, dw, real, noises[1], shifts[12:16], hp)
s = gan.engine.generator_step(real, noises[1], shifts[12:16], update=False)
errs = ['%.1e' % rel_err(x, y.numpy()) for x, y in zip(gan.engine.get_grads(0), ref_g['grads'])]
print('seed %d' % seed, 'B %d L %d C %d nu %d ln %d | gen_loss %.6g vs %.6g | G grads' % (B, L, C, nu, ln, s[4], ref_g['gen_loss']), ' '.join(errs))
