"""Assemble the measured sections of profiles/r2_final.md from the artefacts of tools/final_run.sh / scale_run.sh
(gpurun_out/<tag>_*). Prose is kept in profiles/r2_notes.md and prepended.
  python tools/make_r2_profile.py r2 > profiles/r2_final.md"""
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
G = 'gpurun_out'


def load(name):
  try:
    return json.loads(open(os.path.join(G, name)).read().strip().splitlines()[-1])
  except Exception:
    return None


out = []
notes = os.path.join('profiles', 'r2_notes.md')
if os.path.exists(notes):
  out.append(open(notes).read().rstrip() + '\n')

d = load('%s_bench_paper.json' % tag)
if d:
  out.append('## bench.py, paper config, 1 x B200 (`%s_bench_paper.json`)\n' % tag)
  out.append('| | |\n|---|---|')
  out.append('| full WGAN-GP step (device-resident) | %.3f ms = %.0f samples/s (%.0f TFLOP/s effective) |' % (d['ms_per_step'], d['value'], d['config']['tflops_effective']))
  out.append('| end to end, device-resident dataset cache (%d B H2D per step) | %.0f samples/s |' % (d['e2e']['h2d_bytes_per_step'], d['e2e']['value']))
  out.append('| end to end, streaming every batch from pinned host memory (%.0f MB H2D per step) | %.0f samples/s |' % (d['e2e_streaming']['h2d_bytes_per_step'] / 1e6, d['e2e_streaming']['value']))
  out.append('| kernel launches per step | %d |' % (d['gpu_launches'] / d['steps']))
  out.append('| clocks during the timed region | %s |' % d['clocks'])
  r = d['roofline']
  out.append('| conv GEMM kernels (roofline) | %.0f TFLOP/s = %.3f of %.1f (%s), %.3f ms per step |' % (r['achieved'], r['frac'], r['peak'], r['peak_kind'], r['ms_per_step_in_kernel']))
  if d.get('cpu_baseline'):
    out.append('| CPU baseline | %.2f samples/s on %d cores (%s) |' % (d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['kind']))
  out.append('\nLive per-kernel table (`kernels`; CUDA events around every launch of two extra steps -- bracketing adds a few us per launch):\n')
  out.append('| kernel | bound | launches/step | ms/step | achieved | fraction of measured peak |\n|---|---|---:|---:|---:|---:|')
  for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1].get('ms_per_step', 0)):
    if k.startswith('_'):
      continue
    ach = ('%.0f TFLOP/s' % v['achieved_tflops']) if 'achieved_tflops' in v else (('%.0f GB/s' % v['achieved_gbs']) if 'achieved_gbs' in v else '')
    out.append('| `%s` | %s | %.0f | %.3f | %s | %s |' % (k, v['bound'], v['launches_per_step'], v['ms_per_step'], ach, ('%.2f' % v['frac']) if 'frac' in v else ''))
  out.append('')

p = os.path.join(G, '%s_launches.csv' % tag)
if os.path.exists(p):
  out.append('## Launch list of one full step (`ncu --metrics gpu__time_duration.sum --clock-control none`, second step; `launches_%s.csv`)\n' % tag)
  out.append(subprocess.run([sys.executable, 'tools/launch_table.py', p], capture_output=True, text=True).stdout)

ntag = sys.argv[3] if len(sys.argv) > 3 else tag   # the 27-launch ncu --set full capture may carry an earlier tag
p = os.path.join(G, '%s_full.ncu-rep' % ntag)
if os.path.exists(p):
  raw = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  hdr = rows[0]
  names = {'k': 'Kernel Name', 'grid': 'Grid Size', 'dur': 'gpu__time_duration.sum', 'cyc': 'sm__cycles_elapsed.max',
           'ops': 'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
           'dr': 'dram__bytes_read.sum', 'dw': 'dram__bytes_write.sum', 'dthr': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'l2hit': 'lts__t_sector_hit_rate.pct', 'regs': 'launch__registers_per_thread', 'l2sm': 'lts__t_sectors_srcunit_tex_op_read.sum'}
  idx = {k: hdr.index(v) for k, v in names.items() if v in hdr}

  def f(r, k):
    try:
      return float(r[idx[k]].replace(',', ''))
    except Exception:
      return float('nan')
  out.append('## `ncu --set full --clock-control none --import-source on -k regex:rsgemm3_tc|wgrad2_tc|ghead_tc -s 27 -c 27` (one critic sub-step + the next generator layers)\n')
  out.append('tensor % = `sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed` (issued tcgen05 FLOPs per elapsed cycle against the pipe peak).\n')
  out.append('| # | kernel | us | SM cycles | tensor % of peak (elapsed) | DRAM rd MB | DRAM wr MB | DRAM thr % | L2->SM MB | L2 hit % | regs |\n|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|')
  gem_bytes, gem_n = 0.0, 0
  for i, r in enumerate(rows[2:]):
    k = r[idx['k']].split('(')[0].replace('void ', '')
    out.append('| %d | `%s` | %.1f | %.0fk | %.1f | %.1f | %.1f | %.1f | %.0f | %.0f | %s |' % (
        i, k, f(r, 'dur'), f(r, 'cyc') / 1e3, f(r, 'ops'), f(r, 'dr'), f(r, 'dw'), f(r, 'dthr'), f(r, 'l2sm') * 32 / 1e6, f(r, 'l2hit'), r[idx['regs']]))
    if 'rsgemm' in k:
      gem_bytes += (f(r, 'dr') + f(r, 'dw')) * 1e6
      gem_n += 1
  out.append('')
  if gem_n:
    commit = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
    json.dump({'conv_gemm_dram_bytes_per_launch': gem_bytes / gem_n, 'launches': gem_n, 'commit': commit,
               'source': 'ncu --set full, %s_full.ncu-rep (dram__bytes_read.sum + dram__bytes_write.sum)' % ntag},
              open('profiles/r2_ncu_summary.json', 'w'), indent=1)

p = os.path.join(G, '%s_wgrad.ncu-rep' % tag)
if os.path.exists(p):
  raw = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  hdr = rows[0]
  c = {k: hdr.index(v) for k, v in {'k': 'Kernel Name', 'grid': 'Grid Size', 'dur': 'gpu__time_duration.sum', 'cyc': 'sm__cycles_elapsed.max',
       'ops': 'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
       'dr': 'dram__bytes_read.sum', 'dw': 'dram__bytes_write.sum', 'regs': 'launch__registers_per_thread'}.items()}
  out.append('## Weight-gradient kernel of the FINAL build (`wgrad2p_tc_kernel`, CTA pairs + bulk reduce-add epilogue): `ncu --set full -k regex:wgrad2p?_tc -c 6`\n')
  out.append('The 27-launch table above was captured one build earlier (single-CTA `wgrad2_tc_kernel` with the bulk epilogue, rows 18 - 22); '
             'these are the same five launches (critic conv5 ... conv1) of the final build. Before either change the five read 57.1 / 69.6 / 67.6 / 53.7 / 62.4 %.\n')
  out.append('| # | kernel | grid | us | SM cycles | tensor % of peak (elapsed) | DRAM rd MB | DRAM wr MB | regs |\n|---|---|---|---:|---:|---:|---:|---:|---:|')
  for i, r in enumerate(rows[2:]):
    out.append('| %d | `%s` | %s | %.1f | %.0fk | %.1f | %.1f | %.1f | %s |' % (i, r[c['k']].split('(')[0], r[c['grid']], float(r[c['dur']]), float(r[c['cyc']].replace(',', '')) / 1e3,
               float(r[c['ops']]), float(r[c['dr']]), float(r[c['dw']]), r[c['regs']]))
  out.append('')

p = os.path.join(G, '%s_layers.txt' % tag)
if os.path.exists(p):
  out.append('## Per-layer GEMM kernels in isolation (`tools/bench_layers.py --iters 30`, CUDA events, back to back = power-capped clocks)\n\n```')
  out.append(''.join(l for l in open(p) if 'conv' in l).rstrip())
  out.append('```\n')

d = load('%s_bench_gp.json' % tag)
if d:
  out.append('## Gradient-penalty-only microbench (BASELINE configs[4]; `bench.py --config gp`, all four passes)\n')
  out.append('| batch | ms | samples/s | TFLOP/s | of sustained bf16 | e2e samples/s (xhat from pinned host, GP read back) |\n|---:|---:|---:|---:|---:|---:|')
  for b, v in d['config']['sweep'].items():
    out.append('| %s | %.3f | %.0f | %.0f | %.2f | %.0f |' % (b, v['ms'], v['samples_per_s'], v['tflops'], v['frac_of_sustained_bf16'], v['e2e_samples_per_s']))
  out.append('')

d = load('%s_bench_scaled.json' % tag)
if d:
  out.append('## Scaled model (BASELINE configs[3]: num_units 128, 8192 x 512, batch 64 per GPU; `bench.py --config scaled`, 1 GPU)\n')
  out.append('%.2f ms per step = %.1f samples/s = %.0f TFLOP/s effective; conv GEMM kernels %.0f TFLOP/s (%.2f of sustained); e2e (cache) %.1f, streaming %.1f samples/s.\n'
             % (d['ms_per_step'], d['value'], d['config']['tflops_effective'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'], d['e2e_streaming']['value']))

d = load('%s_bench_fp32_b16.json' % tag)
if d:
  out.append('## fp32 CUDA-core path, BASELINE configs[0] shape (batch 16, fp32) on the GPU\n')
  out.append('%.2f ms per step = %.0f samples/s (the reference-precision path used for the 1e-4 parity tests; CPU oracle port on the same shape: see `cpu_baseline`).\n' % (d['ms_per_step'], d['value']))

stag = sys.argv[2] if len(sys.argv) > 2 else tag   # scaling runs may carry an earlier tag (python tools/make_r2_profile.py r2d r2)
for st in ('%sc_scale' % stag, '%sa_scale' % stag, '%sb_scale' % stag):
  rows = [load('%s_n%d.json' % (st, n)) for n in (1, 2, 4, 8)]
  if all(rows):
    what = {'c': 'final build: overlap + NCCL above 2 ranks, peer memory at 2 (= profiles/r2_scale_n*.json)', 'a': 'earlier build of this round: overlap + NCCL at every size',
            'b': 'earlier build: ONE-SHOT peer-memory exchange at every size (7 x 16 MB pulled per rank at 8 GPUs)'}[st[len(stag)]]
    out.append('## Scaling on one 8-GPU box (`tools/scale_run.sh`, 128 samples per GPU) -- %s\n' % what)
    out.append('| GPUs | ms/step | samples/s | efficiency | e2e (cache) samples/s | efficiency | e2e streaming |\n|---:|---:|---:|---:|---:|---:|---:|')
    for n, r in zip((1, 2, 4, 8), rows):
      out.append('| %d | %.3f | %.0f | %.3f | %.0f | %.3f | %.0f |' % (n, r['ms_per_step'], r['value'], r['value'] / (n * rows[0]['value']),
                 r['e2e']['value'], r['e2e']['value'] / (n * rows[0]['e2e']['value']), r['e2e_streaming']['value']))
    out.append('')
print('\n'.join(out))
