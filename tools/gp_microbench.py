"""BASELINE.json config 5: gradient-penalty-only microbench (critic forward + double-backward GP terms),
batch sweep at seq 2048 x 102.  Times cg_critic_step with update disabled minus nothing else: the GP path is the
x-hat third of the concatenated batch, so this reports the full critic sub-step and the GP-only debug tap."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_hparams
from calciumgan_b200.models.registry import get_models

batches = [int(x) for x in sys.argv[1:]] or [32, 64, 128, 256]
for B in batches:
  hp = make_hparams(B)
  g, d = get_models(hp, None)
  eng = g.engine
  x = torch.rand(B, 2048, 102, device='cuda')
  sh = [3, -2, 5, 0]
  for _ in range(3):
    eng.gp_debug(x, sh)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  n = 10
  e0.record()
  for _ in range(n):
    eng.gp_debug(x, sh)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / n
  gf = 2 * 1.378 * B   # critic forward + input-gradient chain (2 x D fwd FLOPs); the wgrad / lin-fwd passes are in cg_critic_step
  print('GP forward+dgrad  B=%4d  %7.3f ms  %8.1f samples/s  %7.1f TFLOP/s' % (B, ms, B / ms * 1e3, gf / ms))
  eng.close()
