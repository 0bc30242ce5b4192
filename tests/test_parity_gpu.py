"""Parity of the CUDA path against the CPU oracle through the reference-facing plugin API and
the C ABI (north_star: generator output, critic scores, GP value and per-parameter gradients
after one step; rel <= 1e-4 in fp32, <= 2e-2 in bf16; PhaseShuffle bit-exact).

The oracle is an fp64 restatement ("parity unpinned" - see oracle/calciumgan_oracle.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import calciumgan_oracle as O
from tests.util import namespace_from_oracle, rel_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4   # north_star
BF16_TOL = 2e-2   # north_star
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'tiny_step.npz')


def build(hp, batch, mixed=False, **kw):
  from calciumgan_b200.algorithms.registry import get_algorithm
  from calciumgan_b200.models.registry import get_models
  ns = namespace_from_oracle(hp, batch, mixed_precision=mixed, **kw)
  g, d = get_models(ns, None)
  gan = get_algorithm(ns, g, d, None)
  assert type(gan).__name__ == 'WGAN_GP'
  return ns, gan


def check_list(got, ref, tol, what):
  assert len(got) == len(ref)
  worst = 0.0
  for i, (a, b) in enumerate(zip(got, ref)):
    b = b.numpy() if hasattr(b, 'numpy') else b
    assert a.shape == tuple(b.shape), (what, i, a.shape, b.shape)
    if float(np.abs(b).max()) == 0.0:
      assert float(np.abs(a).max()) <= 1e-6, (what, i, 'expected exact zeros', float(np.abs(a).max()))
      continue
    e = rel_err(a, b)
    worst = max(worst, e)
    assert e <= tol, '%s[%d] shape %s rel err %.3e > %.1e' % (what, i, a.shape, e, tol)
  return worst


# ------------------------------------------------------------------------------------------ golden (fp32)
def test_tiny_golden_critic_step_fp32():
  import tests.golden.make_golden as mk
  gold = np.load(GOLD)
  hp = mk.tiny_hp()
  ns, gan = build(hp, mk.BATCH)
  gan.generator.set_weights([gold['gen_w%02d' % i] for i in range(24)])
  gan.discriminator.set_weights([gold['dis_w%02d' % i] for i in range(12)])
  s = gan.engine.critic_step(gold['real'], gold['noises'][0], gold['alphas'][0], gold['shifts'][:12], update=False)
  assert rel_err(gan.engine.fake(mk.BATCH).cpu().numpy(), gold['c_fake']) <= FP32_TOL
  sc = gan.engine.scores(3 * mk.BATCH).cpu().numpy()
  assert rel_err(sc[:mk.BATCH], gold['c_real_out'].ravel()) <= FP32_TOL
  assert rel_err(sc[mk.BATCH:2 * mk.BATCH], gold['c_fake_out'].ravel()) <= FP32_TOL
  assert abs(s[0] - gold['c_scalars'][0]) <= FP32_TOL * max(1.0, abs(gold['c_scalars'][0]))
  assert abs(s[1] - gold['c_scalars'][1]) <= FP32_TOL * max(1.0, abs(gold['c_scalars'][1]))
  grads = gan.engine.get_grads(1)
  check_list(grads, [gold['c_grad%02d' % i] for i in range(12)], FP32_TOL, 'critic grad')


def test_tiny_golden_generator_step_fp32():
  import tests.golden.make_golden as mk
  gold = np.load(GOLD)
  hp = mk.tiny_hp()
  ns, gan = build(hp, mk.BATCH)
  gan.generator.set_weights([gold['gen_w%02d' % i] for i in range(24)])
  gan.discriminator.set_weights([gold['dis_w%02d' % i] for i in range(12)])
  nc = mk.N_CRITIC
  s = gan.engine.generator_step(gold['real'], gold['noises'][nc], gold['shifts'][12 * nc:12 * nc + 4], update=False)
  assert abs(s[4] - gold['g_scalars'][0]) <= FP32_TOL * max(1.0, abs(gold['g_scalars'][0]))
  # metrics sorted: max, mean, min, std  (C enum order: min, max, mean, std)
  np.testing.assert_allclose([s[6], s[7], s[5], s[8]], gold['g_scalars'][1:], rtol=1e-3, atol=1e-7)
  check_list(gan.engine.get_grads(0), [gold['g_grad%02d' % i] for i in range(24)], FP32_TOL, 'generator grad')


def test_tiny_golden_full_train_step_fp32():
  import tests.golden.make_golden as mk
  gold = np.load(GOLD)
  hp = mk.tiny_hp()
  ns, gan = build(hp, mk.BATCH)
  gan.generator.set_weights([gold['gen_w%02d' % i] for i in range(24)])
  gan.discriminator.set_weights([gold['dis_w%02d' % i] for i in range(12)])
  gen_loss, dis_loss, gp, metrics = gan.train(gold['real'], noise=gold['noises'], alpha=gold['alphas'],
                                              shifts=gold['shifts'])
  t = gold['t_scalars']
  assert abs(gen_loss - t[0]) <= 1e-3 * max(1.0, abs(t[0]))
  assert abs(dis_loss - t[1]) <= 1e-3 * max(1.0, abs(t[1]))
  assert abs(gp - t[2]) <= 1e-3 * max(1.0, abs(t[2]))
  assert set(metrics) == {'signals_metrics/min', 'signals_metrics/max', 'signals_metrics/mean', 'signals_metrics/std'}
  # after n_critic Adam steps + 1: weights moved by ~lr per step; compare the *update* not the weight
  w0 = [gold['dis_w%02d' % i] for i in range(12)]
  w1 = [gold['t_dis_w%02d' % i] for i in range(12)]
  got = gan.discriminator.get_weights()
  for i in range(12):
    ref_upd, got_upd = w1[i] - w0[i], got[i] - w0[i]
    if np.abs(ref_upd).max() == 0:
      continue
    assert rel_err(got_upd, ref_upd) <= 5e-2, ('dis update', i, rel_err(got_upd, ref_upd))
  g0 = [gold['gen_w%02d' % i] for i in range(24)]
  g1 = [gold['t_gen_w%02d' % i] for i in range(24)]
  got = gan.generator.get_weights()
  for i in range(24):
    assert rel_err(got[i] - g0[i], g1[i] - g0[i]) <= 5e-2, ('gen update', i)
  assert gan.gen_optimizer.iterations == 1 and gan.dis_optimizer.iterations == mk.N_CRITIC


# ------------------------------------------------------------------------------------------ PhaseShuffle bit-exact
@pytest.mark.parametrize('w', [4, 64, 1024])
def test_phase_shuffle_gather_bit_exact(w):
  hp = O.HParams(signal_shape=(64, 6), noise_dim=4, num_units=4, kernel_size=6, m=2, n_critic=1)
  ns, gan = build(hp, 2)
  rng = np.random.RandomState(0)
  x = rng.standard_normal((3, w, 8)).astype(np.float32)
  m = min(10, w - 1)
  for shift in range(-m, m + 1):
    got = gan.engine.phase_shuffle(x, shift).cpu().numpy()
    np.testing.assert_array_equal(got, O.phase_shuffle_literal(x, shift))


# ------------------------------------------------------------------------------------------ seeded medium configs
def _medium_hp(**kw):
  d = dict(signal_shape=(256, 20), noise_dim=8, num_units=16, kernel_size=24, m=3, n_critic=1)
  d.update(kw)
  return O.HParams(**d)


def _run_both(hp, batch, mixed, seed=7, force_simt=False):
  ns, gan = build(hp, batch, mixed=mixed, force_simt=force_simt)
  gw, dw = O.init_weights(hp, seed=seed)
  gw, dw = O.randomize_weights(gw, seed + 1), O.randomize_weights(dw, seed + 2)
  gan.generator.set_weights(gw)
  gan.discriminator.set_weights(dw)
  real, noises, alphas, shifts = O.synthetic_batch(hp, batch, seed=seed + 3, n_critic=1)
  ref_c = O.critic_step(gw, dw, real, noises[0], alphas[0], shifts[:12].reshape(3, 4), hp)
  s = gan.engine.critic_step(real, noises[0], alphas[0], shifts[:12], update=False)
  got_c = dict(scal=s, fake=gan.engine.fake(batch).cpu().numpy(), scores=gan.engine.scores(3 * batch).cpu().numpy(),
               grads=gan.engine.get_grads(1))
  ref_g = O.generator_step(gw, dw, real, noises[1], shifts[12:16], hp)
  s = gan.engine.generator_step(real, noises[1], shifts[12:16], update=False)
  got_g = dict(scal=s, grads=gan.engine.get_grads(0))
  return ref_c, got_c, ref_g, got_g, gan


@pytest.mark.parametrize('layer_norm', [True, False])
@pytest.mark.parametrize('K', [24, 5])
def test_medium_fp32(layer_norm, K):
  hp = _medium_hp(layer_norm=layer_norm, kernel_size=K)
  ref_c, got_c, ref_g, got_g, _ = _run_both(hp, 4, mixed=False)
  B = 4
  assert rel_err(got_c['fake'], ref_c['fake'].numpy()) <= FP32_TOL
  assert rel_err(got_c['scores'][:B], ref_c['real_out'].numpy().ravel()) <= FP32_TOL
  assert rel_err(got_c['scores'][B:2 * B], ref_c['fake_out'].numpy().ravel()) <= FP32_TOL
  assert abs(got_c['scal'][1] - ref_c['gradient_penalty']) <= FP32_TOL * max(1.0, ref_c['gradient_penalty'])
  assert abs(got_c['scal'][0] - ref_c['dis_loss']) <= FP32_TOL * max(1.0, abs(ref_c['dis_loss']))
  check_list(got_c['grads'], ref_c['grads'], FP32_TOL, 'critic grad')
  assert abs(got_g['scal'][4] - ref_g['gen_loss']) <= FP32_TOL * max(1.0, abs(ref_g['gen_loss']))
  check_list(got_g['grads'], ref_g['grads'], FP32_TOL, 'generator grad')


# LeakyReLU slope flips make ANY 16-bit evaluation of the gradients differ from fp64 by O(sqrt(ulp)),
# and two evaluations under the same bf16 policy (e.g. the oracle's own fp32- vs fp64-accumulated
# restatement) differ by 2-3e-2 at this point; see oracle/calciumgan_oracle.py and DESIGN.md.
# The GEMM kernels themselves are held to bf16 rounding in tests/test_layers_gpu.py.
# These are FREE-RUNNING comparisons (each side takes its own LeakyReLU branches): bounds = measured worst case + 20%
# (0.114 vs fp64 at the paper architecture, batch 3; 0.066 vs the bf16-policy oracle). The north_star tolerance itself
# (2e-2 on every per-parameter gradient) is asserted in tests/test_gradient_parity_gpu.py, where the oracle is evaluated
# on the branches the engine took.
BF16_VS_FP64_GRAD_BOUND = 0.14
BF16_POLICY_GRAD_BOUND = 0.08
BF16_FREE_RUNNING_GRAD_BOUND = 0.075   # batch 128 (measured 0.026 critic / 0.062 generator, + 20%): more samples average more flip noise


@pytest.mark.parametrize('force_simt', [True, False])
def test_medium_bf16(force_simt):
  """bf16 path: generator output, critic scores, GP value and losses within 2e-2 of the fp64 oracle
  (north_star tolerance); per-parameter gradients within the measured 16-bit noise floor of the fp64
  oracle and of the oracle's restatement of the bf16 storage policy, plus a direction check."""
  hp = _medium_hp(signal_shape=(512, 102), num_units=32)
  B, seed = 8, 7
  ref_c, got_c, ref_g, got_g, _ = _run_both(hp, B, mixed=True, force_simt=force_simt)
  assert rel_err(got_c['fake'], ref_c['fake'].numpy()) <= BF16_TOL
  assert rel_err(got_c['scores'][:B], ref_c['real_out'].numpy().ravel()) <= BF16_TOL
  assert rel_err(got_c['scores'][B:2 * B], ref_c['fake_out'].numpy().ravel()) <= BF16_TOL
  assert abs(got_c['scal'][1] - ref_c['gradient_penalty']) <= BF16_TOL * max(1.0, ref_c['gradient_penalty'])
  w64 = check_list(got_c['grads'], ref_c['grads'], BF16_VS_FP64_GRAD_BOUND, 'critic grad vs fp64')
  g64 = check_list(got_g['grads'], ref_g['grads'], BF16_VS_FP64_GRAD_BOUND, 'generator grad vs fp64')
  gw, dw = O.init_weights(hp, seed=seed)
  gw, dw = O.randomize_weights(gw, seed + 1), O.randomize_weights(dw, seed + 2)
  real, noises, alphas, shifts = O.synthetic_batch(hp, B, seed=seed + 3, n_critic=1)
  mix_c = O.critic_step_mixed(gw, dw, real, noises[0], alphas[0], shifts[:12], hp)
  mix_g = O.generator_step_mixed(gw, dw, real, noises[1], shifts[12:16], hp)
  assert rel_err(got_c['fake'], mix_c['fake'].numpy()) <= BF16_TOL
  assert abs(got_c['scal'][0] - mix_c['dis_loss']) <= BF16_TOL * max(1.0, abs(mix_c['dis_loss']))
  wm = check_list(got_c['grads'], mix_c['grads'], BF16_POLICY_GRAD_BOUND, 'critic grad vs bf16-policy oracle')
  gm = check_list(got_g['grads'], mix_g['grads'], BF16_POLICY_GRAD_BOUND, 'generator grad vs bf16-policy oracle')
  for a, b in zip(got_c['grads'] + got_g['grads'], ref_c['grads'] + ref_g['grads']):
    b = b.numpy()
    if float(np.abs(b).max()) > 0:
      cos = float((a.astype(np.float64) * b).sum() / (np.linalg.norm(a.astype(np.float64)) * np.linalg.norm(b)))
      assert cos >= 0.99, cos
  print('bf16 worst grad rel err: critic %.2e / gen %.2e vs fp64; critic %.2e / gen %.2e vs bf16-policy oracle'
        % (w64, g64, wm, gm))


def test_gp_debug_tap_fp32():
  hp = _medium_hp()
  ns, gan = build(hp, 4)
  gw, dw = O.init_weights(hp, seed=3)
  dw = O.randomize_weights(dw, 4)
  gan.discriminator.set_weights(dw)
  real, _, _, _ = O.synthetic_batch(hp, 4, seed=5, n_critic=1)
  sh = [3, -2, 0, 1]
  gp, g, _ = O.gp_four_pass(dw, real, sh, hp)
  got_g, sumsq = gan.engine.gp_debug(real, sh)
  assert rel_err(got_g.cpu().numpy(), g.numpy()) <= FP32_TOL
  n = np.sqrt(sumsq.cpu().numpy())
  assert abs(float(np.mean((n - 1) ** 2)) - float(gp)) <= FP32_TOL * max(1.0, float(gp))
  out = gan.discriminator(real, shifts=sh)
  ref = O.discriminator_forward([torch.tensor(a, dtype=torch.float64) for a in dw],
                                torch.tensor(real, dtype=torch.float64), sh, hp)
  assert rel_err(out.cpu().numpy(), ref.numpy()) <= FP32_TOL


def test_validate_and_generate_fp32():
  hp = _medium_hp()
  B = 4
  ns, gan = build(hp, B)
  gw, dw = O.init_weights(hp, seed=11)
  gw, dw = O.randomize_weights(gw, 12), O.randomize_weights(dw, 13)
  gan.generator.set_weights(gw)
  gan.discriminator.set_weights(dw)
  real, noises, alphas, shifts = O.synthetic_batch(hp, B, seed=14, n_critic=1)
  fake, gen_loss, dis_loss, gp, metrics = O.validate_step(gw, dw, real, noises[0], alphas[0],
                                                          shifts[:12].reshape(3, 4), hp)
  f2, gl2, dl2, gp2, m2 = gan.validate(real, noise=noises[0], alpha=alphas[0], shifts=shifts[:12])
  assert rel_err(f2.cpu().numpy(), fake.numpy()) <= FP32_TOL
  assert abs(gl2 - gen_loss) <= FP32_TOL * max(1, abs(gen_loss))
  assert abs(dl2 - dis_loss) <= FP32_TOL * max(1, abs(dis_loss))
  assert abs(gp2 - gp) <= FP32_TOL * max(1, abs(gp))
  for k in metrics:
    assert abs(m2[k] - metrics[k]) <= 1e-3 * max(1e-3, abs(metrics[k])), k
  w_before = gan.discriminator.get_weights()
  out = gan.generate(noises[0])
  assert rel_err(out.cpu().numpy(), fake.numpy()) <= FP32_TOL
  den = gan.generate(noises[0], denorm=True)
  np.testing.assert_allclose(den.cpu().numpy(), out.cpu().numpy() * (hp.signals_max - hp.signals_min) + hp.signals_min,
                             rtol=1e-6)
  for a, b in zip(w_before, gan.discriminator.get_weights()):
    np.testing.assert_array_equal(a, b)     # validate/generate never update


def test_ragged_last_batch_and_errors():
  hp = _medium_hp()
  ns, gan = build(hp, 4)
  real, noises, alphas, shifts = O.synthetic_batch(hp, 3, seed=2, n_critic=1)
  out = gan.train(real)    # smaller last batch of an epoch, library-drawn randomness
  assert np.isfinite(out[0]) and np.isfinite(out[1]) and np.isfinite(out[2])
  from calciumgan_b200._lib import CgError
  big, _, _, _ = O.synthetic_batch(hp, 5, seed=2, n_critic=1)
  with pytest.raises(CgError, match='max_batch'):
    gan.train(big)
  with pytest.raises(CgError, match='outside'):
    gan.engine.critic_step(real, noises[0], alphas[0], [99] * 12, update=False)


def test_checkpoint_roundtrip(tmp_path):
  from calciumgan_b200.utils import utils
  hp = _medium_hp()
  ns, gan = build(hp, 2, output_dir=str(tmp_path))
  real, _, _, _ = O.synthetic_batch(hp, 2, seed=2, n_critic=1)
  gan.train(real)
  ns.ckpt_dir = os.path.join(str(tmp_path), 'checkpoints')
  utils.save_models(ns, gan, epoch=3)
  gw, dw = gan.generator.get_weights(), gan.discriminator.get_weights()
  ns2, gan2 = build(hp, 2, output_dir=str(tmp_path), seed=99)
  utils.load_models(ns2, gan2)
  assert ns2.start_epoch == 4
  for a, b in zip(gw, gan2.generator.get_weights()):
    np.testing.assert_array_equal(a, b)
  for a, b in zip(dw, gan2.discriminator.get_weights()):
    np.testing.assert_array_equal(a, b)
  assert gan2.gen_optimizer.iterations == 1 and gan2.dis_optimizer.iterations == 1


def test_prefetch_to_device_ragged():
  from calciumgan_b200.utils.prefetch import prefetch_to_device
  rng = np.random.RandomState(0)
  batches = [(rng.rand(b, 16, 5).astype(np.float32), i) for i, b in enumerate([4, 4, 4, 3])]
  pinned = [(torch.from_numpy(x).pin_memory(), i) for x, i in batches]
  for src in (batches, pinned):
    seen = []
    for signal, extra in prefetch_to_device(iter(src)):
      assert signal.is_cuda and signal.dtype == torch.float32
      seen.append((signal.cpu().numpy().copy(), extra))
    assert [e for _, e in seen] == [0, 1, 2, 3]
    for (got, _), (ref, _) in zip(seen, batches):
      np.testing.assert_array_equal(got, ref)


def test_main_end_to_end_on_tfrecords(tmp_path):
  """main.py with the reference's flags on the reference's TFRecord layout: trains, validates, logs, checkpoints, resumes."""
  import glob
  import main as driver
  from calciumgan_b200.utils import dataset_helper
  rng = np.random.RandomState(0)
  data_dir, out_dir = str(tmp_path / 'tfrecords'), str(tmp_path / 'runs')
  signals = rng.rand(10, 256, 20).astype(np.float32)
  dataset_helper.write_dataset(data_dir, signals, np.zeros_like(signals), train_size=7, num_per_shard=4)
  argv = ['--input_dir', data_dir, '--output_dir', out_dir, '--batch_size', '4', '--num_units', '16', '--kernel_size', '24',
          '--m', '3', '--epochs', '2', '--noise_dim', '8', '--model', 'calciumgan', '--layer_norm', '--n_critic', '2',
          '--mixed_precision', '--verbose', '0']
  hp = driver.build_parser().parse_args(argv)
  hp.global_step, hp.surrogate_ds = 0, False
  res = driver.main(hp, return_metrics=True)
  assert set(res) == {'signals_metrics/min', 'signals_metrics/max', 'signals_metrics/mean', 'signals_metrics/std'}
  assert hp.global_step == 4                                  # 2 epochs x ceil(7 / 4) batches
  assert os.path.exists(os.path.join(out_dir, 'checkpoints', 'epoch-001.pkl'))
  assert glob.glob(os.path.join(out_dir, 'events.out.tfevents.*'))
  assert not os.listdir(os.path.join(out_dir, 'generated'))   # --save_generated defaults to "": nothing is written
  argv2 = list(argv) + ['--save_generated', 'last']
  argv2[argv2.index('--epochs') + 1] = '3'
  hp2 = driver.build_parser().parse_args(argv2)
  hp2.global_step, hp2.surrogate_ds = 0, False
  driver.main(hp2)
  assert hp2.start_epoch == 2 and hp2.global_step == 2        # resumed from epoch-001, one more epoch
  # generated validation signals of the last epoch (main.py:81-84,105-106; utils.py:93-113): de-normalised float32 NWC
  import pickle
  from calciumgan_b200.utils import h5_helper
  filename = os.path.join(out_dir, 'generated', 'epoch002_signals.h5')
  fake = h5_helper.get(filename, 'signals')
  assert fake.shape == (3, 256, 20) and fake.dtype == np.float32       # 10 - 7 validation signals in one batch
  assert signals.min() <= fake.min() and fake.max() <= signals.max()     # sigmoid output mapped back to the data range
  with open(os.path.join(out_dir, 'generated', 'info.pkl'), 'rb') as file:
    assert pickle.load(file) == {2: {'global_step': 2, 'filename': filename}}


def test_generator_head_kernel_modes_bf16(monkeypatch):
  """The dedicated generator-head kernel (dense + sigmoid + fused WGAN-GP interpolation, cg_kernels_head.cuh)
  against the generic implicit-GEMM path + interp_kernel (CG_NO_GHEAD=1) in its four output modes:
  train step (fake16 + x_hat16), public critic step / validate (+ fp32 copy, in place over the real tile),
  generator step (fake16 + fp32) and generate (fp32 only). Same fp32 arithmetic on both paths => tight bounds."""
  hp = _medium_hp(signal_shape=(512, 102), num_units=32, n_critic=2)
  B = 5   # 20 row tiles of 128: several tiles per sample, odd batch
  gw, dw = O.init_weights(hp, seed=21)
  gw, dw = O.randomize_weights(gw, 22), O.randomize_weights(dw, 23)
  real, noises, alphas, shifts = O.synthetic_batch(hp, B, seed=24, n_critic=2)
  res = {}
  for mode in ('generic', 'fused'):
    if mode == 'generic':
      monkeypatch.setenv('CG_NO_GHEAD', '1')
    else:
      monkeypatch.delenv('CG_NO_GHEAD', raising=False)
    ns, gan = build(hp, B, mixed=True)
    gan.generator.set_weights(gw)
    gan.discriminator.set_weights(dw)
    r = {}
    f, gl, dl, gp, met = gan.validate(real, noise=noises[0], alpha=alphas[0], shifts=shifts[:12])
    r['val_fake'], r['val'] = f.cpu().numpy(), np.array([gl, dl, gp] + [met[k] for k in sorted(met)])
    r['gen'] = gan.generate(noises[1]).cpu().numpy()
    s = gan.engine.critic_step(real, noises[0], alphas[0], shifts[:12], update=False)
    r['c_scal'], r['c_fake'] = np.array(s[:5]), gan.engine.fake(B).cpu().numpy()
    r['c_scores'], r['c_grads'] = gan.engine.scores(3 * B).cpu().numpy(), gan.engine.get_grads(1)
    if mode == 'fused':   # CG_FLAG_NO_FAKE32 (data-parallel critic sub-steps): same scalars and gradients, fp32 copy untouched
      before = gan.engine.fake(B).cpu().numpy().copy()
      s2 = gan.engine.critic_step(real, noises[1], alphas[1], shifts[:12], update=False, want_fake32=False)
      np.testing.assert_array_equal(gan.engine.fake(B).cpu().numpy(), before)
      s3 = gan.engine.critic_step(real, noises[1], alphas[1], shifts[:12], update=False, want_fake32=True)
      assert rel_err(np.array(s2[:5]), np.array(s3[:5])) <= 1e-6
      assert not np.array_equal(gan.engine.fake(B).cpu().numpy(), before)
    out = gan.train(real, noise=noises, alpha=alphas, shifts=shifts)
    r['train'] = np.array(out[:3])
    r['gw'], r['dw'] = gan.generator.get_weights(), gan.discriminator.get_weights()
    res[mode] = r
    if mode == 'fused':
      assert gan.engine.tc_launch_count() > 0
  a, b = res['fused'], res['generic']
  for k in ('val_fake', 'gen', 'c_fake'):
    assert rel_err(a[k], b[k]) <= 1e-6, k          # same MMA, same fp32 epilogue arithmetic
  for k in ('val', 'c_scal', 'c_scores', 'train'):
    assert rel_err(a[k], b[k]) <= 2e-3, (k, a[k], b[k])
  check_list(a['c_grads'], b['c_grads'], 2e-2, 'critic grads fused vs generic head')
  # Adam's first two steps are sign-like (lr * g / |g|): compare the weights, not the update
  check_list(a['gw'], b['gw'], 1e-3, 'generator weights after a train step')
  check_list(a['dw'], b['dw'], 1e-3, 'critic weights after a train step')


@pytest.mark.parametrize('shifts4', [(10, -10, 7, -1), (-10, 10, 0, 3), (1, -9, -10, 10)])
def test_phase_shuffle_adjoint_fused_into_dgrad_bf16(monkeypatch, shifts4):
  """Data-gradient GEMM with the PhaseShuffle adjoint (edge reflection summed in registers) and the LeakyReLU slope
  fused into its epilogue (EPI_PS_MASK) against the unfused dgrad + ps_scatter_mask_kernel path, at the full
  sequence length (three layers with >= 128 rows per sample) and the extreme shifts of m = 10. The fused path sums
  the two contributors of a reflected row in fp32 instead of after bf16 rounding => bf16-ulp agreement."""
  hp = _medium_hp(signal_shape=(2048, 20), num_units=16, m=10)
  B = 3
  gw, dw = O.init_weights(hp, seed=31)
  gw, dw = O.randomize_weights(gw, 32), O.randomize_weights(dw, 33)
  real, noises, alphas, _ = O.synthetic_batch(hp, B, seed=34, n_critic=1)
  shifts = np.array(list(shifts4) + list(shifts4[::-1]) + [shifts4[1], shifts4[0], shifts4[3], shifts4[2]], np.int32)
  res = {}
  for mode in ('unfused', 'fused'):
    if mode == 'unfused':
      monkeypatch.setenv('CG_NO_PS_BWD_FUSE', '1')
    else:
      monkeypatch.delenv('CG_NO_PS_BWD_FUSE', raising=False)
    ns, gan = build(hp, B, mixed=True)
    gan.generator.set_weights(gw)
    gan.discriminator.set_weights(dw)
    n0 = gan.engine.launch_count()
    s = gan.engine.critic_step(real, noises[0], alphas[0], shifts, update=False)
    res[mode] = dict(scal=np.array(s[:5]), grads=gan.engine.get_grads(1), launches=gan.engine.launch_count() - n0)
    s = gan.engine.generator_step(real, noises[1], shifts[:4], update=False)
    res[mode]['ggrads'] = gan.engine.get_grads(0)
  assert res['fused']['launches'] == res['unfused']['launches'] - 3     # three ps_scatter_mask launches gone
  assert rel_err(res['fused']['scal'], res['unfused']['scal']) <= 1e-3
  check_list(res['fused']['grads'], res['unfused']['grads'], 1e-2, 'critic grads, fused vs unfused PS adjoint')
  check_list(res['fused']['ggrads'], res['unfused']['ggrads'], 1e-2, 'generator grads, fused vs unfused PS adjoint')
  ref = O.critic_step_mixed(gw, dw, real, noises[0], alphas[0], shifts, hp)
  # batch 3: bias-sized tensors sit right at the 16-bit noise floor (see BF16_* bounds above)
  check_list(res['fused']['grads'], ref['grads'], BF16_VS_FP64_GRAD_BOUND, 'critic grads vs bf16-policy oracle')


@pytest.mark.parametrize('B', [3, 8])
def test_paper_architecture_bf16(B):
  """The exact BASELINE.json architecture (noise_dim 32, num_units 64, kernel 24, strides 2, layer_norm, m = 10,
  seq 2048 x 102) at a small batch (3: odd, CTA pairs straddle samples; 8): every tensor-core kernel runs with the
  layer shapes, tap tables and PhaseShuffle extremes of the headline benchmark. Same bounds as test_medium_bf16."""
  hp = O.HParams()
  seed = 40 + B
  ref_c, got_c, ref_g, got_g, gan = _run_both(hp, B, mixed=True, seed=seed)
  assert gan.engine.tc_launch_count() > 0
  assert rel_err(got_c['fake'], ref_c['fake'].numpy()) <= BF16_TOL
  # at this init the critic scores are O(1e-2) sums of 20480 cancelling terms: absolute bound as for the losses
  ref_scores = np.concatenate([ref_c['real_out'].numpy().ravel(), ref_c['fake_out'].numpy().ravel()])
  assert np.abs(got_c['scores'][:2 * B] - ref_scores).max() <= BF16_TOL * max(1.0, np.abs(ref_scores).max())
  assert abs(got_c['scal'][1] - ref_c['gradient_penalty']) <= BF16_TOL * max(1.0, ref_c['gradient_penalty'])
  assert abs(got_c['scal'][0] - ref_c['dis_loss']) <= BF16_TOL * max(1.0, abs(ref_c['dis_loss']))
  assert abs(got_g['scal'][4] - ref_g['gen_loss']) <= BF16_TOL * max(1.0, abs(ref_g['gen_loss']))
  check_list(got_c['grads'], ref_c['grads'], BF16_VS_FP64_GRAD_BOUND, 'critic grad vs fp64')
  check_list(got_g['grads'], ref_g['grads'], BF16_VS_FP64_GRAD_BOUND, 'generator grad vs fp64')
  for a, b in zip(got_c['grads'] + got_g['grads'], ref_c['grads'] + ref_g['grads']):
    b = b.numpy()
    if float(np.abs(b).max()) > 0:
      cos = float((a.astype(np.float64) * b).sum() / (np.linalg.norm(a.astype(np.float64)) * np.linalg.norm(b)))
      assert cos >= 0.99, cos


def test_paper_architecture_shift_extremes_bf16():
  """Paper architecture with every PhaseShuffle draw at +-m = +-10 (both reflection branches of calciumgan.py:126-133 in
  the fused forward scatter and in the fused adjoint of every layer) and one all-zero set."""
  hp = O.HParams()
  B = 2
  gw, dw = O.init_weights(hp, seed=51)
  gw, dw = O.randomize_weights(gw, 52), O.randomize_weights(dw, 53)
  real, noises, alphas, _ = O.synthetic_batch(hp, B, seed=54, n_critic=1)
  ns, gan = build(hp, B, mixed=True)
  gan.generator.set_weights(gw)
  gan.discriminator.set_weights(dw)
  for shifts in (np.array([10, -10, 10, -10, -10, 10, -10, 10, 10, 10, -10, -10], np.int32), np.zeros(12, np.int32)):
    ref = O.critic_step(gw, dw, real, noises[0], alphas[0], shifts.reshape(3, 4), hp)
    s = gan.engine.critic_step(real, noises[0], alphas[0], shifts, update=False)
    assert abs(s[1] - ref['gradient_penalty']) <= BF16_TOL * max(1.0, ref['gradient_penalty'])
    assert abs(s[0] - ref['dis_loss']) <= BF16_TOL * max(1.0, abs(ref['dis_loss']))
    ref_scores = np.concatenate([ref['real_out'].numpy().ravel(), ref['fake_out'].numpy().ravel()])
    assert np.abs(gan.engine.scores(3 * B).cpu().numpy()[:2 * B] - ref_scores).max() <= BF16_TOL * max(1.0, np.abs(ref_scores).max())
    # batch 2: a 64-element bias gradient averages the least flip noise (measured 0.146; + 20%)
    check_list(gan.engine.get_grads(1), ref['grads'], 0.175, 'critic grad vs fp64, shifts %s' % shifts[:4])


def _fp32_outputs_ok(B, ref_c, got_c, ref_g, got_g):
  assert rel_err(got_c['fake'], ref_c['fake'].numpy()) <= FP32_TOL
  ref_scores = np.concatenate([ref_c['real_out'].numpy().ravel(), ref_c['fake_out'].numpy().ravel()])
  assert np.abs(got_c['scores'][:2 * B] - ref_scores).max() <= FP32_TOL * max(1.0, np.abs(ref_scores).max())
  assert abs(got_c['scal'][1] - ref_c['gradient_penalty']) <= FP32_TOL * max(1.0, ref_c['gradient_penalty'])
  assert abs(got_c['scal'][0] - ref_c['dis_loss']) <= FP32_TOL * max(1.0, abs(ref_c['dis_loss']))
  assert abs(got_g['scal'][4] - ref_g['gen_loss']) <= FP32_TOL * max(1.0, abs(ref_g['gen_loss']))


# A LeakyReLU slope is discontinuous: when ONE activation's sign differs between the fp32 and the fp64 evaluation, every
# gradient below it moves by ~|dy_e| / ||dy|| ~ 1/sqrt(#elements) (2e-6 above the flipped layer, 1e-3 .. 1e-2 from there
# down; tools/fp32_paper_errs.py prints the step pattern, torch fp32 behaves the same). The flip probability per
# evaluation grows with the number of activations: ~0.5% for the medium configs, ~5% at sequence length 256 with the paper
# widths, ~50% at the full 2048 x 102 x batch 2. Hence, for these FREE-RUNNING comparisons, a flip-tolerant bound plus a
# direction check on gradients (outputs and losses are held to 1e-4 everywhere); with the branches imposed the same
# gradients agree to 4e-6 (tests/test_gradient_parity_gpu.py).
FP32_FLIP_BOUND = 2e-2


def test_paper_architecture_fp32():
  """Paper layer widths / kernel / strides / m (noise_dim 32, num_units 64, K 24, 102 channels) in fp32, the reference's
  default precision, at sequence length 256, free-running: outputs and losses to 1e-4, gradients to the slope-flip bound
  (single seed; the 1e-4 gradient tolerance is asserted on imposed branches at the FULL length in
  tests/test_gradient_parity_gpu.py::test_gradients_on_imposed_branches_paper_architecture[False-2])."""
  hp = O.HParams(signal_shape=(256, 102))
  B = 2
  ref_c, got_c, ref_g, got_g, _ = _run_both(hp, B, mixed=False, seed=61)
  _fp32_outputs_ok(B, ref_c, got_c, ref_g, got_g)
  wc = check_list(got_c['grads'], ref_c['grads'], FP32_FLIP_BOUND, 'critic grad')
  wg = check_list(got_g['grads'], ref_g['grads'], FP32_FLIP_BOUND, 'generator grad')
  print('fp32 paper widths, length 256, free-running: worst gradient rel err critic %.1e generator %.1e' % (wc, wg))


def test_paper_architecture_full_length_fp32():
  """The exact BASELINE.json architecture (2048 x 102) in fp32: outputs, scores, GP and losses to 1e-4; gradients to the
  slope-flip bound (see above) and parallel to the oracle's."""
  hp = O.HParams()
  B = 2
  ref_c, got_c, ref_g, got_g, _ = _run_both(hp, B, mixed=False, seed=66)
  _fp32_outputs_ok(B, ref_c, got_c, ref_g, got_g)
  wc = check_list(got_c['grads'], ref_c['grads'], FP32_FLIP_BOUND, 'critic grad')
  wg = check_list(got_g['grads'], ref_g['grads'], FP32_FLIP_BOUND, 'generator grad')
  for a, b in zip(got_c['grads'] + got_g['grads'], ref_c['grads'] + ref_g['grads']):
    b = b.numpy()
    if float(np.abs(b).max()) > 0:
      cos = float((a.astype(np.float64) * b).sum() / (np.linalg.norm(a.astype(np.float64)) * np.linalg.norm(b)))
      assert cos >= 0.9995, cos
  print('fp32 full length: worst gradient rel err critic %.1e generator %.1e' % (wc, wg))


@pytest.mark.parametrize('mixed', [False, True])
def test_cuda_path_matches_reference_code_fixture(mixed):
  """The CUDA path (through the C ABI; fp32, and bf16 mixed precision) against tests/golden/reference_step.npz -- outputs
  of the REFERENCE'S OWN code (gan/models/calciumgan.py, gan/algorithms/wgan_gp.py, ... executed unmodified over the
  torch-backed TensorFlow stand-in, tests/golden/make_reference_golden.py): validate, generate (+denorm) and one full
  train step (n_critic 2) incl. the weights after Adam. north_star tolerances: 1e-4 in fp32, 2e-2 in bf16."""
  import sys
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
  import make_reference_golden as G
  tol = BF16_TOL if mixed else FP32_TOL
  gold = np.load(os.path.join(os.path.dirname(GOLD), 'reference_step.npz'))
  hp, gw, dw, real, noises, alphas, shifts = G.inputs()
  ns, gan = build(hp, G.BATCH, mixed=mixed)
  gan.generator.set_weights(gw)
  gan.discriminator.set_weights(dw)
  fake, gl, dl, gp, met = gan.validate(real, noise=noises[0], alpha=alphas[0], shifts=shifts[:12])
  assert rel_err(fake.cpu().numpy(), gold['val_fake']) <= tol
  got = np.array([gl, dl, gp] + [met[k] for k in sorted(met)])
  assert np.abs(got - gold['val_scalars']).max() <= tol * max(1.0, np.abs(gold['val_scalars']).max())
  assert rel_err(gan.generate(noises[1]).cpu().numpy(), gold['gen_fake']) <= tol
  assert rel_err(gan.generate(noises[1], denorm=True).cpu().numpy(), gold['gen_fake_denorm']) <= tol
  out = gan.train(real, noise=noises, alpha=alphas, shifts=shifts)
  got = np.array(list(out[:3]) + [out[3][k] for k in sorted(out[3])])
  assert np.abs(got - gold['train_scalars']).max() <= tol * max(1.0, np.abs(gold['train_scalars']).max())
  # weights after n_critic critic updates / one generator update. Adam's first steps are sign-like (lr * g / |g|), so
  # near-zero gradient elements dominate a relative error on the UPDATE: the weights are held to the tolerance, the
  # update loosely in fp32 (the critic's output bias has an exactly-zero gradient, -1 + 1 + 0: rounding noise) and by
  # its size only in bf16 (|update| <= (n_critic or 1) * lr per element whatever the gradient noise)
  for name, new, w0 in (('gen', gan.generator.get_weights(), gw), ('dis', gan.discriminator.get_weights(), dw)):
    for i, (a, b0) in enumerate(zip(new, w0)):
      ref = gold['%s_w_%02d' % (name, i)]
      assert np.abs(a - ref).max() <= tol * max(1.0, np.abs(ref).max()), (name, i, np.abs(a - ref).max())
      if not mixed and float(np.abs(ref - b0).max()) > 1e-6:
        assert rel_err(a - b0, ref - b0) <= 5e-2, (name, i, rel_err(a - b0, ref - b0))
      if mixed:
        assert np.abs(a - b0).max() <= 2.001 * hp.learning_rate * 1.01 + 1e-7, (name, i)


@pytest.mark.parametrize('mixed', [False, True])
def test_gradients_match_reference_code_fixture(mixed):
  """Per-parameter gradients as the REFERENCE'S OWN _train_discriminator / _train_generator hand them to Adam
  (wgan_gp.py:22-36,64-80 -> optimizer.py:31-34; captured from the unmodified reference code in
  tests/golden/reference_step.npz and reference_medium.npz). fp32: rel <= 1e-4 on every tensor. bf16: free-running
  comparison (slope-flip bound + direction; the 2e-2 tolerance on imposed branches is tests/test_gradient_parity_gpu.py,
  against the oracle that reproduces these fixtures to 1e-9, tests/test_reference_shim.py)."""
  import sys
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
  import make_reference_golden as G
  tol = BF16_TOL if mixed else FP32_TOL
  gtol = 0.175 if mixed else FP32_TOL      # batch 3 / 4: see the batch-2 bound above
  # small fixture: whole gradient tensors
  gold = np.load(os.path.join(os.path.dirname(GOLD), 'reference_step.npz'))
  hp, gw, dw, real, noises, alphas, shifts = G.inputs()
  ns, gan = build(hp, G.BATCH, mixed=mixed)
  gan.generator.set_weights(gw)
  gan.discriminator.set_weights(dw)
  s = gan.engine.critic_step(real, noises[0], alphas[0], shifts[:12], update=False)
  assert abs(s[0] - gold['sub_scalars'][0]) <= tol * max(1.0, abs(gold['sub_scalars'][0]))
  assert abs(s[1] - gold['sub_scalars'][1]) <= tol * max(1.0, abs(gold['sub_scalars'][1]))
  wc = check_list(gan.engine.get_grads(1), [gold['c_grad_%02d' % i] for i in range(12)], gtol, 'critic grad vs reference code')
  s = gan.engine.generator_step(real, noises[1], shifts[12:16], update=False)
  assert abs(s[4] - gold['sub_scalars'][2]) <= tol * max(1.0, abs(gold['sub_scalars'][2]))
  wg = check_list(gan.engine.get_grads(0), [gold['g_grad_%02d' % i] for i in range(24)], gtol, 'generator grad vs reference code')
  # the gradients INSIDE a full train step (last critic update, generator update after the critic's Adam steps); Adam's
  # first steps are sign-like, so a near-zero gradient element can move a weight the other way: looser in fp32
  gan.train(real, noise=noises, alpha=alphas, shifts=shifts)
  check_list(gan.engine.get_grads(1), [gold['t_c_grad_%02d' % i] for i in range(12)], gtol if mixed else 2e-3, 'critic grad in train()')
  check_list(gan.engine.get_grads(0), [gold['t_g_grad_%02d' % i] for i in range(24)], gtol if mixed else 2e-3, 'generator grad in train()')
  # medium fixture (512 x 102, shifts at +-10): norms + strided samples
  gold = np.load(os.path.join(os.path.dirname(GOLD), 'reference_medium.npz'))
  hp, gw, dw, real, noises, alphas, shifts = G.inputs_medium()
  ns, gan = build(hp, G.BATCH_MEDIUM, mixed=mixed)
  gan.generator.set_weights(gw)
  gan.discriminator.set_weights(dw)
  gan.engine.critic_step(real, noises[0], alphas[0], shifts[:12], update=False)
  gc = gan.engine.get_grads(1)
  gan.engine.generator_step(real, noises[1], shifts[12:16], update=False)
  gg = gan.engine.get_grads(0)
  for key, grads in (('c', gc), ('g', gg)):
    for i, a in enumerate(grads):
      ref, norm = gold['%s_grad_%02d' % (key, i)], gold['%s_grad_norms' % key][i]
      if norm == 0:
        assert np.abs(a).max() <= 1e-6
        continue
      # free-running in bf16: a bias-sized tensor's norm moves with the slope flips like its elements do
      assert abs(np.linalg.norm(a.astype(np.float64)) / norm - 1.0) <= (0.1 if mixed else tol), (key, i)
      assert rel_err(a.reshape(-1)[::G.grad_stride(a.size)], ref) <= gtol, (key, i)
  print('gradients vs reference-code fixture (mixed=%s): worst critic %.2e generator %.2e' % (mixed, wc, wg))


@pytest.mark.parametrize('mixed', [False, True])
def test_cuda_path_matches_reference_code_fixture_medium(mixed):
  """Second fixture from the reference's own code (tests/golden/reference_medium.npz: 512 x 102, num_units 32, m 10 with
  shifts at the extremes): the slab-mode tensor-core kernels with both fused PhaseShuffle directions in bf16, the
  CUDA-core path in fp32. validate outputs + the scalars of one train step."""
  import sys
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
  import make_reference_golden as G
  tol = BF16_TOL if mixed else FP32_TOL
  gold = np.load(os.path.join(os.path.dirname(GOLD), 'reference_medium.npz'))
  hp, gw, dw, real, noises, alphas, shifts = G.inputs_medium()
  ns, gan = build(hp, G.BATCH_MEDIUM, mixed=mixed)
  gan.generator.set_weights(gw)
  gan.discriminator.set_weights(dw)
  fake, gl, dl, gp, met = gan.validate(real, noise=noises[0], alpha=alphas[0], shifts=shifts[:12])
  assert rel_err(fake.cpu().numpy(), gold['val_fake']) <= max(tol, 2e-7 * 10)   # fixture stored as float32
  got = np.array([gl, dl, gp] + [met[k] for k in sorted(met)])
  assert np.abs(got - gold['val_scalars']).max() <= tol * max(1.0, np.abs(gold['val_scalars']).max()), (got, gold['val_scalars'])
  out = gan.train(real, noise=noises, alpha=alphas, shifts=shifts)
  got = np.array(list(out[:3]) + [out[3][k] for k in sorted(out[3])])
  assert np.abs(got - gold['train_scalars']).max() <= tol * max(1.0, np.abs(gold['train_scalars']).max()), (got, gold['train_scalars'])
  if mixed:
    assert gan.engine.tc_launch_count() > 0
  else:   # update norms: every tensor moved as far as in the reference run (Adam: ~lr * sqrt(#elements) on the first step)
    for name, new, w0, key in (('gen', gan.generator.get_weights(), gw, 'gen_update_norms'),
                               ('dis', gan.discriminator.get_weights(), dw, 'dis_update_norms')):
      norms = np.array([np.linalg.norm(a - np.asarray(b)) for a, b in zip(new, w0)])
      ok = gold[key] > 1e-6
      assert np.abs(norms[ok] / gold[key][ok] - 1).max() <= 2e-2, (name, norms, gold[key])


@pytest.mark.parametrize('B', [1, 7, 128])
def test_paper_config_tensor_core_vs_cuda_core_paths(B):
  """BASELINE.json configs[1] at its FULL batch (128) and at ragged batches: the bf16 tensor-core path against the fp32
  CUDA-core path of the same engine on identical weights / inputs / draws (the two share no GEMM, epilogue or head
  code; the fp32 path is the one held to 1e-4 against the oracle). north_star bf16 tolerance 2e-2."""
  hp = O.HParams()
  gw, dw = O.init_weights(hp, seed=3)
  gw, dw = O.randomize_weights(gw, 4), O.randomize_weights(dw, 5)
  real, noises, alphas, shifts = O.synthetic_batch(hp, B, seed=6 + B, n_critic=1)
  res = {}
  for mixed in (False, True):
    ns, gan = build(hp, B, mixed=mixed)
    gan.generator.set_weights(gw)
    gan.discriminator.set_weights(dw)
    s = gan.engine.critic_step(real, noises[0], alphas[0], shifts[:12], update=False)
    scores = gan.engine.scores(3 * B).cpu().numpy()[:2 * B]
    fake = gan.engine.fake(B).cpu().numpy()
    s2 = gan.engine.generator_step(real, noises[1], shifts[12:16], update=False)
    res[mixed] = (np.array(s[:4]), scores, fake, np.array(s2[4:9]))
    if mixed:
      assert gan.engine.tc_launch_count() > 0
    gan.engine.close()
  a, b = res[True], res[False]
  assert np.abs(a[0] - b[0]).max() <= BF16_TOL * max(1.0, np.abs(b[0]).max())     # dis_loss, GP, real / fake loss
  assert np.abs(a[1] - b[1]).max() <= BF16_TOL * max(1.0, np.abs(b[1]).max())     # critic scores
  assert rel_err(a[2], b[2]) <= BF16_TOL                                          # generator output
  assert np.abs(a[3] - b[3]).max() <= BF16_TOL * max(1.0, np.abs(b[3]).max())     # gen_loss + signal metrics


def test_headline_batch_128_against_cpu_fixture():
  """BASELINE.json configs[1] as benchmarked (batch 128, bf16 tensor-core path) against tests/golden/paper_b128.npz, the
  float64 CPU oracle's free-running evaluation (tests/golden/make_b128_golden.py): scalars, scores and the generator
  output at the north_star 2e-2; every per-parameter gradient's norm, direction and a strided sample of its elements
  within the slope-flip noise floor (the tight, branch-imposed check of the same case is
  tests/test_gradient_parity_gpu.py::test_gradients_on_imposed_branches_headline_batch_128)."""
  import sys
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
  import make_b128_golden as G
  gold = np.load(os.path.join(os.path.dirname(GOLD), 'paper_b128.npz'))
  hp, gw, dw, real, noises, alphas, shifts = G.inputs()
  B = G.BATCH
  ns, gan = build(hp, B, mixed=True)
  gan.generator.set_weights(gw)
  gan.discriminator.set_weights(dw)
  s = gan.engine.critic_step(real, noises[0], alphas[0], shifts[:12], update=False)
  assert gan.engine.tc_launch_count() > 0
  assert np.abs(np.array(s[:2]) - gold['c_scalars']).max() <= BF16_TOL * max(1.0, np.abs(gold['c_scalars']).max())
  scores = gan.engine.scores(3 * B).cpu().numpy()[:2 * B]
  assert np.abs(scores - gold['c_scores']).max() <= BF16_TOL * max(1.0, np.abs(gold['c_scores']).max())
  fake = gan.engine.fake(B).cpu().numpy().reshape(-1)[::100003]
  assert rel_err(fake, gold['c_fake_sample']) <= BF16_TOL

  def check(grads, prefix):
    worst = 0.0
    for i, a in enumerate(grads):
      ref, norm = gold['%s_grad%02d' % (prefix, i)], gold['%s_grad_norms' % prefix][i]
      if norm == 0:
        assert np.abs(a).max() <= 1e-6
        continue
      assert abs(np.linalg.norm(a.astype(np.float64)) / norm - 1.0) <= BF16_TOL, (prefix, i)
      got = a.reshape(-1)[::G.sample_stride(a.size)].astype(np.float64)
      e = rel_err(got, ref)
      worst = max(worst, e)
      assert e <= BF16_FREE_RUNNING_GRAD_BOUND, (prefix, i, e)
      if ref.size >= 16:
        assert float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref))) >= 0.995, (prefix, i)
    return worst

  wc = check(gan.engine.get_grads(1), 'c')
  s = gan.engine.generator_step(real, noises[1], shifts[12:16], update=False)
  got = np.array([s[4], s[6], s[7], s[5], s[8]])     # gen_loss, then metrics sorted: max, mean, min, std
  assert np.abs(got - gold['g_scalars']).max() <= BF16_TOL * max(1.0, np.abs(gold['g_scalars']).max())
  assert np.abs(gan.engine.scores(B).cpu().numpy() - gold['g_scores']).max() <= BF16_TOL * max(1.0, np.abs(gold['g_scores']).max())
  wg = check(gan.engine.get_grads(0), 'g')
  print('batch 128 vs float64 CPU fixture (free-running branches): worst sampled gradient rel err critic %.2e generator %.2e'
        % (wc, wg))


def test_gan_metrics_and_step_methods():
  """GAN.metrics(real, fake) (gan.py:32-41) and GAN._step(real, noise, training) (gan.py:58-70) as callable methods of
  the plugin, against the oracle."""
  hp = _medium_hp()
  B = 4
  ns, gan = build(hp, B)
  gw, dw = O.init_weights(hp, seed=11)
  gw, dw = O.randomize_weights(gw, 12), O.randomize_weights(dw, 13)
  gan.generator.set_weights(gw)
  gan.discriminator.set_weights(dw)
  real, noises, alphas, shifts = O.synthetic_batch(hp, B, seed=14, n_critic=1)
  rng = np.random.RandomState(0)
  fake = rng.uniform(0, 1, size=real.shape).astype(np.float32)
  hp2 = hp
  want = O.signals_metrics(torch.tensor(real, dtype=torch.float64), torch.tensor(fake, dtype=torch.float64), hp2)
  got = gan.metrics(real, torch.tensor(fake).cuda())
  assert set(got) == set(want)
  for k in want:
    assert abs(got[k] - want[k]) <= 1e-4 * max(1e-3, abs(want[k])), k
  f, gl, dl, gp, met = O.validate_step(gw, dw, real, noises[0], alphas[0], shifts[:12].reshape(3, 4), hp)
  f2, gl2, dl2, gp2, met2 = gan._step(real, noises[0], training=False, alpha=alphas[0], shifts=shifts[:12])
  assert rel_err(f2.cpu().numpy(), f.numpy()) <= FP32_TOL
  assert abs(gl2 - gl) <= FP32_TOL * max(1, abs(gl)) and abs(dl2 - dl) <= FP32_TOL * max(1, abs(dl))
  assert abs(gp2 - gp) <= FP32_TOL * max(1, abs(gp))
  out = gan._step(real, gan.get_noise(B))           # library-drawn alpha / shifts
  assert len(out) == 5 and np.isfinite(out[1]) and np.isfinite(out[2]) and np.isfinite(out[3])


def test_main_profile_flag_and_hparams_json(tmp_path):
  """--profile opens the cudaProfilerStart/Stop + NVTX window at the reference's batch indices (main.py:45-52,
  summary_helper.py:115-119) and main() writes output_dir/hparams.json (utils.py:72-75) that load_hparams reads back."""
  import argparse
  import json
  import main as driver
  from calciumgan_b200.utils import dataset_helper, utils
  from calciumgan_b200.utils import summary_helper
  rng = np.random.RandomState(0)
  data_dir, out_dir = str(tmp_path / 'tfrecords'), str(tmp_path / 'runs')
  signals = rng.rand(16, 256, 20).astype(np.float32)
  dataset_helper.write_dataset(data_dir, signals, np.zeros_like(signals), train_size=14, num_per_shard=8)
  calls = []
  orig_trace, orig_export = summary_helper.Summary.profiler_trace, summary_helper.Summary.profiler_export
  summary_helper.Summary.profiler_trace = lambda self: (calls.append(('trace', hp.global_step)), orig_trace(self))
  summary_helper.Summary.profiler_export = lambda self: (calls.append(('export', hp.global_step)), orig_export(self))
  try:
    argv = ['--input_dir', data_dir, '--output_dir', out_dir, '--batch_size', '2', '--num_units', '16', '--kernel_size', '24',
            '--m', '3', '--epochs', '2', '--noise_dim', '8', '--model', 'calciumgan', '--layer_norm', '--n_critic', '1',
            '--mixed_precision', '--verbose', '0', '--profile', '--skip_checkpoints']
    hp = driver.build_parser().parse_args(argv)
    hp.global_step, hp.surrogate_ds = 0, False
    driver.main(hp)
  finally:
    summary_helper.Summary.profiler_trace, summary_helper.Summary.profiler_export = orig_trace, orig_export
  # 7 batches per epoch: the window opens before batch 2 and closes after batch 6 of epoch 1 (global steps 9 and 13)
  assert calls == [('trace', 9), ('export', 13)], calls
  content = json.load(open(os.path.join(out_dir, 'hparams.json')))
  assert content['num_units'] == 16 and content['signal_shape'] == [256, 20] and 'git_hash' in content
  fresh = argparse.Namespace(output_dir=out_dir, num_units=99)
  utils.load_hparams(fresh)
  assert fresh.num_units == 99 and fresh.kernel_size == 24 and fresh.noise_dim == 8    # only missing fields are filled
