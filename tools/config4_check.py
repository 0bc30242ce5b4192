"""BASELINE.json config 4 sanity: num_units 128, seq 8192 x 512, batch 64 per GPU - one full step, timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.calciumgan_oracle import HParams
from tests.util import namespace_from_oracle
from calciumgan_b200.algorithms.registry import get_algorithm
from calciumgan_b200.models.registry import get_models
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hp = namespace_from_oracle(HParams(signal_shape=(8192, 512), num_units=128), B, mixed_precision=True)
g, d = get_models(hp, None)
gan = get_algorithm(hp, g, d, None)
print('params G %d D %d, device bytes %.1f GB' % (g.count_params(), d.count_params(), gan.engine.device_bytes() / 1e9))
real = torch.rand(B, 8192, 512, device='cuda')
for _ in range(2):
  out = gan.train(real)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
  out = gan.train(real)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print('config4: B=%d  %.1f ms/step  %.1f samples/s  %.1f TFLOP/s (1882 GF/sample)  losses %s  tc %d/%d launches' %
      (B, ms, B / ms * 1e3, 1882 * B / ms, out[:3], gan.engine.tc_launch_count(), gan.engine.launch_count()))
