"""bench.py — WGAN-GP samples/sec on synthetic (batch, 2048, 102) signals (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one full WGAN-GP step: n_critic = 5 critic updates + 1 generator update, including Adam
(and, for N > 1, the NCCL gradient all-reduces), through the reference-facing plugin API
(`get_models` / `get_algorithm` / `gan.train`).  N = 1 workload = BASELINE.json configs[1]: CalciumGAN
paper config, batch 128, seq 2048 x 102, bf16 mixed precision.  N > 1: the same per GPU (weak scaling).

`--impl reference` times the reference's algorithm on the host CPU: TensorFlow 2.3.1 is not installable
in this image, so it is the oracle port (oracle/calciumgan_oracle.py, torch CPU fp32), on a bounded
sample (BASELINE.json configs[0]: batch 16) of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = 'WGAN-GP samples/sec (seq 2048x102 neurons)'
UNIT = 'samples/s'
# algorithmic FLOPs of one full step per sample, paper config (SURVEY §8d / BASELINE.md §3)
GF_PER_SAMPLE_STEP = 87.86


def measured_peaks():
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
      p = json.load(f)
    return p, 'measured'
  except Exception:
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler(object):
  """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
  Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, index=0):
    self.rows, self.proc, self.index = [], None, index

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '20'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None

  def stop_all(self):
    sm, mx = [], []
    for r in self.rows:
      try:
        sm.append(float(r[1]))
        mx.append(float(r[2]))
      except Exception:
        pass
    sm.sort()
    return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': [],
            'samples': len(sm)}

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([x.strip() for x in line.split(',')])

  def stop(self, t0=None, t1=None):
    """Summarise the samples whose timestamp lies in the timed window [t0, t1] (host epoch seconds)."""
    import datetime
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for r in self.rows:
      try:
        ts = datetime.datetime.strptime(r[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
        if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.05):
          continue
        sm.append(float(r[1]))
        mx.append(float(r[2]))
        for n, v in zip(names, r[4:8]):
          if v.lower().startswith('active'):
            reasons.add(n)
      except Exception:
        pass
    if not sm and t0 is not None:   # sampler slower than the timed window: fall back to every sample of this run
      out = self.stop_all()
      out['window'] = 'whole run (no sample fell inside the timed window)'
      return out
    sm.sort()
    return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'reasons': sorted(reasons), 'samples': len(sm)}


def make_hparams(batch, mixed=True):
  from oracle.calciumgan_oracle import HParams   # dataclass of defaults only (no arithmetic)
  from tests.util import namespace_from_oracle
  return namespace_from_oracle(HParams(), batch, mixed_precision=mixed)


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_step_time(steps, warmup, batch=16, budget_s=150.0):
  """The oracle port of the reference step on the host cores (fp32, torch CPU)."""
  import numpy as np
  import torch
  from oracle import calciumgan_oracle as O
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  hp = O.HParams()
  gw, dw = O.init_weights(hp, seed=1234)
  st = O.TrainState.create(gw, dw, dtype=torch.float32)
  real, noises, alphas, shifts = O.synthetic_batch(hp, batch, seed=1234)
  times, t_start = [], time.time()
  for i in range(warmup + steps):
    t0 = time.time()
    O.train_step(st, real, noises, alphas, shifts, hp, dtype=torch.float32)
    dt = time.time() - t0
    if i >= warmup:
      times.append(dt)
    if time.time() - t_start > budget_s and len(times) >= 1:
      break
  return float(np.mean(times)), len(times), cores, batch


def run_reference(args, rank):
  if rank != 0:
    return
  warm = min(args.warmup, 1)
  sec, done, cores, batch = cpu_step_time(args.steps, warm)
  v = batch / sec
  sample = 'full WGAN-GP step (5 critic + 1 generator update, Adam) at batch %d fp32, %d timed steps' % (batch, done)
  print(json.dumps({
      'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': done,
      'warmup': warm, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': 'CalciumGAN paper config WGAN-GP step, seq 2048x102; CPU arm on a bounded sample: batch %d' % batch,
                 'note': 'oracle port (torch CPU) of the reference algorithm; TensorFlow 2.3.1 is not installable here'},
      'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
      'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }))


# ------------------------------------------------------------------------------------------ GPU arm
def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--batch', type=int, default=128, help='per-GPU batch (BASELINE.json: 128)')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--fp32', action='store_true', help='debug: fp32 CUDA-core path (not the headline config)')
  args = ap.parse_args()
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))

  if args.impl == 'reference':
    run_reference(args, rank)
    return

  import numpy as np
  import torch
  import torch.distributed as dist
  assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
  torch.cuda.set_device(local_rank)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
  from calciumgan_b200.algorithms.registry import get_algorithm
  from calciumgan_b200.models.registry import get_models

  clocks = ClockSampler(local_rank)
  if rank == 0:
    clocks.start()     # nvidia-smi needs ~1 s to start streaming: launch it before the engine is built
  warmup = max(args.warmup, 3)
  B = args.batch
  hparams = make_hparams(B, mixed=not args.fp32)
  generator, discriminator = get_models(hparams, None)
  gan = get_algorithm(hparams, generator, discriminator, None)
  eng = gan.engine

  rng = np.random.RandomState(1234 + rank)
  real_host = torch.from_numpy(rng.uniform(0, 1, size=(B, 2048, 102)).astype(np.float32)).pin_memory()
  real_dev = real_host.cuda(non_blocking=True)
  torch.cuda.synchronize()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
      t = torch.tensor([ms], device='cuda')
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t.item())
    return ms

  # ---- device-resident arm: inputs already in HBM when the timed region starts
  def step_resident():
    return gan.train(real_dev)

  for _ in range(warmup):
    out = step_resident()
  l0 = eng.launch_count()
  t_wall0 = time.time()
  ms = timed(step_resident, args.steps)
  t_wall1 = time.time()
  launches = eng.launch_count() - l0
  clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
  ms_per_step = ms / args.steps
  value = world * B * args.steps / (ms * 1e-3)

  # ---- end-to-end arm: host (pinned) buffers in, host floats out, copies inside the timed region.
  # Every step's batch is copied from pinned host memory; the copy of step i+1 is double-buffered on a side stream
  # (calciumgan_b200.utils.prefetch, the stand-in for the reference's tf.data prefetch) and gan.train returns Python
  # floats, i.e. the scalars are read back from the device every step.
  from calciumgan_b200.utils.prefetch import prefetch_to_device
  host_batches = [real_host, real_host.clone().pin_memory()]

  def run_e2e(steps):
    for signal, _ in prefetch_to_device((host_batches[i % 2] for i in range(steps))):
      gan.train(signal)

  run_e2e(2)
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  run_e2e(args.steps)
  e1.record()
  barrier()
  ms_e2e = e0.elapsed_time(e1)
  if world > 1:
    t = torch.tensor([ms_e2e], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
  e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
  h2d = real_host.numel() * 4
  d2h = 16 * 4

  # ---- live roofline of the dominant kernel (tcgen05 implicit-GEMM conv), CUDA events on its stream
  peaks, peak_kind = measured_peaks()
  roof, kernels = None, None
  eng.profile(True)          # every rank runs these steps (they contain the DP all-reduces)
  for _ in range(2):
    step_resident()
  rep = eng.profile_report()
  eng.profile(False)
  barrier()
  if rank == 0:
    g, w = rep['gemm'], rep['wgrad']
    peak_tf = float(peaks.get('bf16_tflops_sustained', peaks['bf16_tflops']))
    ach = g['flops'] / (g['ms'] * 1e-3) / 1e12 if g['ms'] > 0 else 0.0
    roof = {'bound': 'tensor',
            'kernel': 'rsgemm3_tc_kernel / rsgemm_tc_kernel (tcgen05 cta_group::2 implicit-GEMM conv, conv-transpose, '
                      'data gradients, GP linearised forward, per-timestep dense)',
            'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf,
            'peak_kind': peak_kind + ' sustained cuBLAS bf16',
            # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the launches captured with
            # `ncu --set full` in profiles/r1_final.md (tensor-bound kernel: informative only)
            'traffic': 63.7e6, 'traffic_source': 'profiles/r1_final.md',
            'launches_per_step': g['launches'] / 2, 'ms_per_step_in_kernel': g['ms'] / 2}
    ach_w = w['flops'] / (w['ms'] * 1e-3) / 1e12 if w['ms'] > 0 else 0.0
    kernels = {'wgrad2_tc_kernel': {'achieved_tflops': ach_w, 'frac': ach_w / peak_tf, 'ms_per_step': w['ms'] / 2,
                                   'launches_per_step': w['launches'] / 2},
               'gemm_share_of_step': (g['ms'] + w['ms']) / 2 / ms_per_step}
    hd = rep.get('head')
    if hd and hd['launches']:
      # HBM-bound generator head (dense + sigmoid + fused interpolation), algorithmic bytes per step:
      # critic sub-steps read activations (rows x 128 bf16) + the real batch (rows x 102 fp32) and write fake + x_hat
      # (2 x rows x 128 bf16); the generator step reads activations and writes fake bf16 + fake fp32.
      rows = B * 2048
      nc = hparams.n_critic
      step_bytes = nc * rows * (128 * 2 + 102 * 4 + 2 * 128 * 2) + rows * (128 * 2 + 128 * 2 + 102 * 4)
      hbm = float(peaks.get('hbm_gbs', 6650.0))
      gbs = step_bytes * 2 / (hd['ms'] * 1e-3) / 1e9
      kernels['ghead_tc_kernel'] = {'bound': 'hbm', 'achieved_gbs': gbs, 'peak_gbs': hbm, 'frac': gbs / hbm,
                                    'ms_per_step': hd['ms'] / 2, 'launches_per_step': hd['launches'] / 2,
                                    'bytes_per_step': step_bytes}

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    sec, done, cores, cb = cpu_step_time(1, 0)
    cpu = {'value': cb / sec, 'unit': UNIT, 'cores': cores, 'kind': 'port',
           'sample': 'one full WGAN-GP step (5 critic + 1 generator update, Adam) at batch %d fp32 on the host, '
                     'oracle port in torch CPU (TensorFlow 2.3.1 not installable): %.1f s' % (cb, sec)}

  if rank == 0:
    print(json.dumps({
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32' if args.fp32 else 'bf16', 'data': 'synthetic',
        'config': {'workload': 'CalciumGAN paper config (noise_dim 32, num_units 64, kernel 24, strides 2, layer_norm, '
                               'm=10) WGAN-GP, n_critic 5 + 1 generator update, batch %d per GPU, seq 2048 x 102' % B,
                   'global_batch': world * B, 'parallelism': 'dp%d' % world,
                   'l2': 'working set per step (~1.5 GB of activations) exceeds the 126 MB L2; no explicit flush',
                   'tflops_effective': value * GF_PER_SAMPLE_STEP / 1e3},
        'clocks': clk, 'gpu_launches': int(launches),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': ms_e2e / args.steps},
        'roofline': roof, 'kernels': kernels, 'cpu_baseline': cpu,
        'last_losses': {'gen': out[0], 'dis': out[1], 'gp': out[2]},
    }))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
