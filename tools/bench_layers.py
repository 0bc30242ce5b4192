"""Per-layer GEMM kernel timing through cg_bench_layer (CUDA events on the engine stream)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import calciumgan_oracle as O
from tests.util import namespace_from_oracle
from calciumgan_b200.models.registry import get_models

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=128)
ap.add_argument('--iters', type=int, default=20)
ap.add_argument('--simt', action='store_true')
ap.add_argument('--fp32', action='store_true')
ap.add_argument('--only', default='', help='comma list such as D1fwd,D2dgrad,G5fwd')
a = ap.parse_args()
hp = O.HParams()
ns = namespace_from_oracle(hp, a.batch, mixed_precision=not a.fp32, force_simt=a.simt)
g, d = get_models(ns, None)
eng = g.engine
peak = 1665.4
try:
  peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['bf16_tflops']
except Exception:
  pass
rows = []
for which, name, batches in ((1, 'D conv', 3 * a.batch), (0, 'G convT', a.batch)):
  for layer in range(1, 6):
    for pass_, pn in ((0, 'fwd'), (1, 'dgrad'), (2, 'wgrad')):
      if which == 0 and pass_ != 0:
        continue
      B = batches
      if a.only and ('%s%d%s' % (name[0], layer, pn)) not in a.only.split(','):
        continue
      try:
        ms, fl = eng.bench_layer(which, layer, pass_, B, a.iters)
      except Exception as e:
        print('%-8s L%d %-6s B=%4d  failed: %s' % (name, layer, pn, B, str(e)[:60]))
        continue
      tf = fl / ms / 1e9
      rows.append((name, layer, pn, B, ms, tf))
      print('%-8s L%d %-6s B=%4d  %8.3f ms  %8.1f TFLOP/s  %5.1f%% of %.0f' % (name, layer, pn, B, ms, tf, 100 * tf / peak, peak))
print('tc launches', eng.tc_launch_count(), 'of', eng.launch_count())
