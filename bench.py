"""bench.py — WGAN-GP samples/sec on synthetic (batch, 2048, 102) signals (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config paper|scaled|gp]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one full WGAN-GP step: n_critic = 5 critic updates + 1 generator update, including Adam
(and, for N > 1, the NCCL gradient all-reduces), through the reference-facing plugin API
(`get_models` / `get_algorithm` / `gan.train`). Workloads (BASELINE.json `configs`):

  --config paper  (default) configs[1] / [2]: CalciumGAN paper config, batch 128 per GPU, seq 2048 x 102, bf16 mixed
                  precision; N > 1 = batch-sharded data parallel, 128 per GPU (weak scaling).
  --config scaled configs[3]: num_units 128, seq 8192 x 512, batch 64 per GPU (quoted on 8 GPUs; runs on any N).
  --config gp     configs[4]: gradient-penalty-only microbench -- critic forward + double-backward GP (all four passes:
                  forward, data-gradient chain, linearised forward, weight gradients), batch sweep 32..1024 at
                  seq 2048 x 102; `value` is the batch-128 point, `config.sweep` holds every batch.

JSON keys beyond the contract: `roofline` (tcgen05 conv GEMMs, live CUDA-event time on the engine stream against
MEASURED_PEAKS.json), `kernels` (every other kernel of the step: tensor or HBM roofline from algorithmic work / live event
time), `cpu_baseline`, `e2e` (+ `e2e_streaming`), `clocks`.

End-to-end arms. `e2e`: the epoch loop a user runs (main.py): host dataset -> device-resident cache (the reference's
`train_ds.cache()`, gan/utils/dataset_helper.py:171, kept in HBM) -> shuffled index vector copied host->device every step
-> gather kernel -> `gan.train` -> Python floats (device->host read of the scalars every step). The cache is filled
during warm-up by streaming every batch from pinned host memory; the timed region then moves 8 bytes per sample.
`e2e_streaming`: no cache, every step's batch copied from pinned host memory (double-buffered under the previous step).

`--impl reference` times the reference's algorithm on the host CPU: TensorFlow 2.3.1 is not installable in this image,
so it is the oracle port (oracle/calciumgan_oracle.py, torch CPU fp32), on a bounded sample (BASELINE.json configs[0]:
batch 16) of the same workload. The GPU arm imports nothing from oracle/ or tests/.
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
# algorithmic GFLOP per real sample (SURVEY 8d / BASELINE.md 3): full step, and the four gradient-penalty passes alone
WORKLOADS = {
    'paper': dict(seq=2048, channels=102, num_units=64, batch=128, gf_step=87.86, gf_gp=5.511,
                  name='CalciumGAN paper config (noise_dim 32, num_units 64, kernel 24, strides 2, layer_norm, m=10)'),
    'scaled': dict(seq=8192, channels=512, num_units=128, batch=64, gf_step=1882.0, gf_gp=119.2,
                   name='scaled model (num_units 128, kernel 24, strides 2, layer_norm, m=10), seq 8192 x 512'),
}


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


def make_hparams(batch, mixed=True, config='paper'):
  """The hparams Namespace main.py builds for the README.md:92 command line (parsed by main.build_parser(), so the CLI
  surface is exercised) plus the fields gan/utils/dataset_helper.py:84-91,120-136 fills in from the dataset."""
  import main as driver
  w = WORKLOADS[config]
  argv = ['--batch_size', str(batch), '--noise_dim', '32', '--num_units', str(w['num_units']), '--kernel_size', '24',
          '--strides', '2', '--m', '10', '--layer_norm', '--algorithm', 'wgan-gp', '--model', 'calciumgan', '--n_critic', '5',
          '--gradient_penalty', '10.0', '--learning_rate', '0.0001', '--verbose', '0']
  if mixed:
    argv.append('--mixed_precision')
  hp = driver.build_parser().parse_args(argv)
  hp.signal_shape, hp.num_channels = (w['seq'], w['channels']), w['channels']
  hp.sequence_length, hp.num_neurons = w['seq'], w['channels']
  hp.noise_shape = (hp.noise_dim,)
  hp.normalize, hp.fft, hp.conv2d = True, False, False
  hp.signals_min, hp.signals_max = 0.0, 1.0
  hp.global_step, hp.surrogate_ds = 0, False
  return hp


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


# ------------------------------------------------------------------------------------------ GPU arms
def kernel_table(table, steps, peaks, peak_kind):
  """bench `roofline` + `kernels` from the engine's live per-kernel table (CUDA events around every launch of `steps`
  profiled steps; event bracketing adds a few microseconds to each launch, so short kernels read pessimistic)."""
  peak_tf = float(peaks.get('bf16_tflops_sustained', peaks['bf16_tflops']))
  hbm = float(peaks.get('hbm_gbs', 6650.0))
  kernels, total_ms = {}, 0.0
  for name, r in table.items():
    ms = r['ms'] / steps
    total_ms += ms
    e = {'bound': r['bound'], 'ms_per_step': ms, 'launches_per_step': r['launches'] / steps}
    if r['bound'] == 'tensor' and r['ms'] > 0:
      e['achieved_tflops'] = r['flops'] / (r['ms'] * 1e-3) / 1e12
      e['frac'] = e['achieved_tflops'] / peak_tf
    elif r['bytes'] > 0 and r['ms'] > 0:
      e['achieved_gbs'] = r['bytes'] / (r['ms'] * 1e-3) / 1e9
      e['peak_gbs'], e['frac'] = hbm, e['achieved_gbs'] / hbm
      e['bytes_per_step'] = r['bytes'] / steps
    kernels[name] = e
  g = table.get('rsgemm_tc')
  roof = None
  if g and g['ms'] > 0:
    ach = g['flops'] / (g['ms'] * 1e-3) / 1e12
    traffic, src = None, None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
      with open(os.path.join(ROOT, 'profiles', 'r2_ncu_summary.json')) as f:
        t = json.load(f)
      traffic, src = t['conv_gemm_dram_bytes_per_launch'], 'profiles/r2_ncu_summary.json (static ncu capture, commit %s)' % t.get('commit')
    except Exception:
      pass
    roof = {'bound': 'tensor',
            'kernel': 'rsgemm3_tc_kernel / rsgemm_tc_kernel (tcgen05 cta_group::2 implicit-GEMM conv, conv-transpose, '
                      'data gradients, GP linearised forward, per-timestep dense)',
            'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf,
            'peak_kind': peak_kind + ' sustained cuBLAS bf16', 'traffic': traffic, 'traffic_source': src,
            'launches_per_step': g['launches'] / steps, 'ms_per_step_in_kernel': g['ms'] / steps}
  return roof, kernels, total_ms


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--config', default='paper', choices=['paper', 'scaled', 'gp'])
  ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default: the BASELINE.json batch of the config)')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-dp-overlap', action='store_true', help='A/B: do not run the next generator forward under the all-reduce')
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
  if args.config == 'gp':
    run_gp(args, rank, world)
    if world > 1:
      dist.destroy_process_group()
    return
  from calciumgan_b200.algorithms.registry import get_algorithm
  from calciumgan_b200.models.registry import get_models

  clocks = ClockSampler(local_rank)
  if rank == 0:
    clocks.start()     # nvidia-smi needs ~1 s to start streaming: launch it before the engine is built
  warmup = max(args.warmup, 3)
  W = WORKLOADS[args.config]
  B = args.batch or W['batch']
  L_, C_ = W['seq'], W['channels']
  hparams = make_hparams(B, mixed=not args.fp32, config=args.config)
  generator, discriminator = get_models(hparams, None)
  gan = get_algorithm(hparams, generator, discriminator, None)
  gan.no_dp_overlap = args.no_dp_overlap
  eng = gan.engine

  rng = np.random.RandomState(1234 + rank)
  n_cached = 4 if args.config == 'paper' else 2          # host "dataset": a few batches of this rank's shard
  dataset = torch.from_numpy(rng.uniform(0, 1, size=(n_cached * B, L_, C_)).astype(np.float32)).pin_memory()
  real_dev = dataset[:B].cuda(non_blocking=True)
  torch.cuda.synchronize()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
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
  out = None

  def resident(steps):
    nonlocal out
    for _ in range(steps):
      out = gan.train(real_dev)

  resident(warmup)
  l0 = eng.launch_count()
  t_wall0 = time.time()
  ms = timed(lambda: resident(args.steps))
  t_wall1 = time.time()
  launches = eng.launch_count() - l0
  clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
  ms_per_step = ms / args.steps
  value = world * B * args.steps / (ms * 1e-3)

  # ---- end-to-end arm 1 (headline e2e): epoch loop over a host dataset with the device-resident cache
  from calciumgan_b200.utils.dataset_cache import DeviceDatasetCache
  from calciumgan_b200.utils.prefetch import prefetch_to_device
  cache = DeviceDatasetCache(eng, n_cached * B, (L_, C_))
  host_batches = [dataset[i * B:(i + 1) * B] for i in range(n_cached)]
  for signal, _ in cache.fill_from(iter(host_batches)):       # first epoch (warm-up): stream + cache every batch
    gan.train(signal)
  shuffle_rng = np.random.RandomState(99 + rank)

  def epochs_cached(steps):
    done = 0
    while done < steps:
      for signal, _ in cache.batches(B, shuffle=True, rng=shuffle_rng, drop_remainder=True):
        gan.train(signal)                                        # returns Python floats: scalars read back every step
        done += 1
        if done == steps:
          break

  epochs_cached(2)
  h2d0 = cache.h2d_bytes
  ms_e2e = timed(lambda: epochs_cached(args.steps))
  e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
  h2d_cached = (cache.h2d_bytes - h2d0) / args.steps
  d2h = 16 * 4

  # ---- end-to-end arm 2: no cache, every step's batch streamed from pinned host memory (double-buffered)
  def streaming(steps):
    for signal, _ in prefetch_to_device((host_batches[i % n_cached] for i in range(steps))):
      gan.train(signal)

  streaming(2)
  ms_stream = timed(lambda: streaming(args.steps))
  stream_value = world * B * args.steps / (ms_stream * 1e-3)

  # ---- live per-kernel table: CUDA events on the engine stream around every launch of two more steps
  peaks, peak_kind = measured_peaks()
  eng.profile(True)          # every rank runs these steps (they contain the DP all-reduces)
  resident(2)
  table = eng.profile_table()
  eng.profile(False)
  barrier()
  roof = kernels = None
  if rank == 0:
    roof, kernels, in_kernels = kernel_table(table, 2, peaks, peak_kind)
    tensor_ms = sum(k['ms_per_step'] for k in kernels.values() if k['bound'] == 'tensor')
    kernels['_summary'] = {'tensor_kernel_share_of_step': tensor_ms / ms_per_step,
                           'other_kernel_share_of_step': (in_kernels - tensor_ms) / ms_per_step,
                           'note': 'event-bracketed launches; shares are upper bounds (bracketing defeats programmatic dependent launch)'}

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    sec, done, cores, cb = cpu_step_time(1, 0)
    cpu = {'value': cb / sec, 'unit': UNIT, 'cores': cores, 'kind': 'port',
           'sample': 'one full WGAN-GP step (5 critic + 1 generator update, Adam) at batch %d fp32, seq 2048 x 102, on the '
                     'host, oracle port in torch CPU (TensorFlow 2.3.1 not installable): %.1f s' % (cb, sec)}

  if rank == 0:
    metric = METRIC if args.config == 'paper' else 'WGAN-GP samples/sec (seq %dx%d neurons)' % (L_, C_)
    print(json.dumps({
        'metric': metric, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32' if args.fp32 else 'bf16', 'data': 'synthetic',
        'config': {'workload': '%s WGAN-GP, n_critic 5 + 1 generator update, batch %d per GPU, seq %d x %d'
                               % (W['name'], B, L_, C_),
                   'baseline_config': 'configs[%d]' % ({'paper': 1 if world == 1 else 2, 'scaled': 3}[args.config]),
                   'global_batch': world * B, 'parallelism': 'dp%d' % world,
                   'l2': 'working set per step (>= 1.5 GB of activations) exceeds the 126 MB L2; no explicit flush',
                   'tflops_effective': value * W['gf_step'] / 1e3,
                   'dp_overlap': (not args.no_dp_overlap) if world > 1 else None},
        'clocks': clk, 'gpu_launches': int(launches),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d_cached), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': ms_e2e / args.steps,
                'path': 'host dataset -> device-resident cache (filled during warm-up: %d MB streamed once) -> per step: '
                        'shuffled indices H2D, gather kernel, gan.train, scalars D2H' % (n_cached * B * L_ * C_ * 4 >> 20)},
        'e2e_streaming': {'value': stream_value, 'unit': UNIT, 'h2d_bytes_per_step': int(B * L_ * C_ * 4),
                          'd2h_bytes_per_step': int(d2h), 'ms_per_step': ms_stream / args.steps,
                          'path': 'no cache: every batch copied from pinned host memory, double-buffered'},
        'roofline': roof, 'kernels': kernels, 'cpu_baseline': cpu,
        'last_losses': {'gen': out[0], 'dis': out[1], 'gp': out[2]},
    }))
  if world > 1:
    dist.destroy_process_group()


def run_gp(args, rank, world):
  """BASELINE.json configs[4]: the gradient penalty alone (cg_gp_gradient: forward, data-gradient chain, linearised
  forward, weight gradients), batch sweep at seq 2048 x 102. Every rank runs the sweep on its own GPU (no exchange step:
  replicas), rank 0 reports rank-0 numbers times the world size."""
  import numpy as np
  import torch
  import torch.distributed as dist
  from calciumgan_b200.models.registry import get_models
  W = WORKLOADS['paper']
  batches = [32, 64, 128, 256, 512, 1024]
  hparams = make_hparams(max(batches), mixed=not args.fp32)
  generator, discriminator = get_models(hparams, None)
  eng = generator.engine
  rng = np.random.RandomState(7 + rank)
  xhat = torch.from_numpy(rng.uniform(0, 1, size=(max(batches), W['seq'], W['channels'])).astype(np.float32)).cuda()
  shifts = rng.randint(-10, 11, size=(64, 4))
  warmup = max(args.warmup, 3)
  peaks, peak_kind = measured_peaks()
  peak_tf = float(peaks.get('bf16_tflops_sustained', peaks['bf16_tflops']))
  sweep, l0 = {}, eng.launch_count()
  for B in batches:
    x = xhat[:B]
    for i in range(warmup):
      eng.gp_gradient(x, shifts[i % 64], sync=False)
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
      eng.gp_gradient(x, shifts[i % 64], sync=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # end to end: xhat from pinned host memory every call, GP value read back
    host = xhat[:B].cpu().pin_memory()
    t0 = time.time()
    for i in range(args.steps):
      gp = eng.gp_gradient(host.cuda(non_blocking=True), shifts[i % 64], sync=True)
    ms_e2e = (time.time() - t0) * 1e3 / args.steps
    tf = B * W['gf_gp'] / ms          # GFLOP per millisecond = TFLOP/s
    sweep[str(B)] = {'samples_per_s': world * B / (ms * 1e-3), 'ms': ms, 'tflops': tf, 'frac_of_sustained_bf16': tf / peak_tf,
                     'e2e_samples_per_s': world * B / (ms_e2e * 1e-3), 'gp': gp}
  if rank == 0:
    ref = sweep['128']
    print(json.dumps({
        'metric': 'gradient-penalty samples/sec (critic forward + double-backward GP, seq 2048x102 neurons)',
        'value': ref['samples_per_s'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
        'ms_per_step': ref['ms'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32' if args.fp32 else 'bf16', 'data': 'synthetic',
        'config': {'workload': 'BASELINE.json configs[4]: gradient-penalty-only microbench, critic of the paper config, all four '
                               'passes (forward, dD/dxhat, linearised forward, dGP/dW), batch sweep 32-1024, seq 2048 x 102; '
                               'value = batch 128', 'parallelism': 'replicas x%d (no exchange step)' % world,
                   'gflop_per_sample': W['gf_gp'], 'sweep': sweep},
        'gpu_launches': int(eng.launch_count() - l0),
        'e2e': {'value': ref['e2e_samples_per_s'], 'unit': UNIT, 'h2d_bytes_per_step': 128 * W['seq'] * W['channels'] * 4,
                'd2h_bytes_per_step': 64},
        'roofline': {'bound': 'tensor', 'kernel': 'the four gradient-penalty passes (rsgemm3_tc / wgrad2_tc + glue), whole call',
                     'achieved': ref['tflops'], 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ref['tflops'] / peak_tf,
                     'peak_kind': peak_kind + ' sustained cuBLAS bf16', 'traffic': None},
        'cpu_baseline': None,
    }))


if __name__ == '__main__':
  main()
