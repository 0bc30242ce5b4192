"""Two full WGAN-GP steps at the paper config (for ncu): the second one is the profiled one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import make_hparams
from calciumgan_b200.algorithms.registry import get_algorithm
from calciumgan_b200.models.registry import get_models

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hp = make_hparams(B)
g, d = get_models(hp, None)
gan = get_algorithm(hp, g, d, None)
real = torch.from_numpy(np.random.RandomState(0).uniform(0, 1, (B, 2048, 102)).astype(np.float32)).cuda()
for _ in range(steps):
  out = gan.train(real)
torch.cuda.synchronize()
print('launches', gan.engine.launch_count(), 'losses', out[:3])
