"""Per-launch GEMM table of one full step (CUDA events around every implicit-GEMM launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['CG_PROF_DUMP'] = '1'
import numpy as np, torch
from bench import make_hparams
from calciumgan_b200.algorithms.registry import get_algorithm
from calciumgan_b200.models.registry import get_models
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
hp = make_hparams(B)
g, d = get_models(hp, None)
gan = get_algorithm(hp, g, d, None)
real = torch.from_numpy(np.random.RandomState(0).uniform(0, 1, (B, 2048, 102)).astype(np.float32)).cuda()
for _ in range(3):
  gan.train(real)
# one critic sub-step + one generator step, profiled
gan.engine.profile(True)
gan.engine.critic_step(real)
gan.engine.generator_step(real)
print(gan.engine.profile_report())
