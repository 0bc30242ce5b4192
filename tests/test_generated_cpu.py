"""`save_fake_signals` and its dataset store (gan/utils/utils.py:50-63,93-113, gan/utils/h5_helper.py,
main.py:81-84): append semantics, selectors, bookkeeping in info.pkl, de-normalisation, and the epochs at which main.py
saves. Runs on whichever backend the image offers (h5py when importable, else the .npy-parts fallback)."""
import argparse
import os
import pickle

import numpy as np
import pytest

from calciumgan_b200.utils import h5_helper, utils


def test_write_appends_along_axis_0_and_selectors(tmp_path):
  filename = str(tmp_path / 'x.h5')
  rng = np.random.RandomState(0)
  a, b = rng.rand(3, 16, 5).astype(np.float32), rng.rand(2, 16, 5).astype(np.float32)
  assert not os.path.exists(filename) and not os.path.exists(h5_helper.parts_dir(filename))
  h5_helper.write(filename, {'signals': a})
  h5_helper.write(filename, {'signals': b, 'spikes': b.astype(np.int8)})
  full = np.concatenate([a, b])
  got = h5_helper.get(filename, 'signals')
  assert got.dtype == np.float32
  np.testing.assert_array_equal(got, full)
  assert h5_helper.get_dataset_length(filename, 'signals') == 5 and h5_helper.get_dataset_length(filename, 'spikes') == 2
  assert h5_helper.contains(filename, 'spikes') and not h5_helper.contains(filename, 'nothing')
  np.testing.assert_array_equal(h5_helper.get(filename, 'signals', neuron=4), full[:, :, 4])
  for t in (0, 2, 3, 4, -1):       # both blocks and the boundary between them
    np.testing.assert_array_equal(h5_helper.get(filename, 'signals', trial=t), full[t])
  with pytest.raises(KeyError):
    h5_helper.get(filename, 'nothing')
  with pytest.raises(AssertionError):
    h5_helper.get(filename, 'signals', neuron=0, trial=0)
  h5_helper.overwrite(filename, 'signals', a[:1])
  np.testing.assert_array_equal(h5_helper.get(filename, 'signals'), a[:1])
  with pytest.raises(KeyError):
    h5_helper.overwrite(filename, 'nothing', a)


def test_append_rejects_a_different_row_shape(tmp_path):
  if h5_helper.backend() == 'h5py':
    pytest.skip('h5py raises its own error type')
  filename = str(tmp_path / 'x.h5')
  h5_helper.write(filename, {'signals': np.zeros((2, 8, 3), np.float32)})
  with pytest.raises(ValueError):
    h5_helper.write(filename, {'signals': np.zeros((2, 8, 4), np.float32)})
  with pytest.raises(ValueError):
    h5_helper.write(filename, {'signals': np.zeros((2, 8, 3), np.float64)})


def _hparams(tmp_path, **kw):
  hp = argparse.Namespace(generated_dir=str(tmp_path), normalize=True, signals_min=-2.0, signals_max=6.0, fft=False,
                          conv2d=False, global_step=17, epochs=25, save_generated='all')
  hp.__dict__.update(kw)
  return hp


def test_save_fake_signals_denormalises_appends_and_records_the_epoch_once(tmp_path):
  hp = _hparams(tmp_path)
  rng = np.random.RandomState(1)
  batches = [rng.rand(4, 32, 6).astype(np.float32), rng.rand(3, 32, 6).astype(np.float32)]    # ragged last batch
  utils.save_fake_signals(hp, 3, batches[0])
  hp.global_step = 99          # the epoch's entry keeps the step at which it was first written (utils.py:110-113)
  utils.save_fake_signals(hp, 3, batches[1])
  filename = os.path.join(str(tmp_path), 'epoch003_signals.h5')
  got = h5_helper.get(filename, 'signals')
  assert got.dtype == np.float32 and got.shape == (7, 32, 6)
  np.testing.assert_allclose(got, np.concatenate(batches) * 8.0 - 2.0, rtol=1e-6)
  utils.save_fake_signals(hp, 10, batches[1])
  with open(os.path.join(str(tmp_path), 'info.pkl'), 'rb') as file:
    info = pickle.load(file)
  assert info == {3: {'global_step': 17, 'filename': filename},
                  10: {'global_step': 99, 'filename': os.path.join(str(tmp_path), 'epoch010_signals.h5')}}


def test_save_fake_signals_accepts_tensors_and_unnormalised_data(tmp_path):
  torch = pytest.importorskip('torch')
  hp = _hparams(tmp_path, normalize=False)
  x = torch.rand(2, 8, 3, dtype=torch.float64, requires_grad=True)
  utils.save_fake_signals(hp, 0, x)
  got = h5_helper.get(os.path.join(str(tmp_path), 'epoch000_signals.h5'), 'signals')
  assert got.dtype == np.float32
  np.testing.assert_allclose(got, x.detach().numpy().astype(np.float32))
  for flag in ('fft', 'conv2d'):     # out of scope: loud, not silently wrong
    with pytest.raises(NotImplementedError):
      utils.save_fake_signals(_hparams(tmp_path, **{flag: True}), 0, x)


def test_epochs_at_which_main_saves_generated_signals(tmp_path):
  """main.py:81-84: 'all' = every 10th epoch and the last, 'last' = only the last, '' = never"""
  hp = _hparams(tmp_path)
  assert [e for e in range(25) if utils.save_generated_at(hp, e)] == [0, 10, 20, 24]
  hp.save_generated = 'last'
  assert [e for e in range(25) if utils.save_generated_at(hp, e)] == [24]
  hp.save_generated = ''
  assert not any(utils.save_generated_at(hp, e) for e in range(25))


def test_dataset_info_sets_the_generated_dir(tmp_path):
  """gan/utils/dataset_helper.py:139-141"""
  from calciumgan_b200.utils import dataset_helper
  data_dir = str(tmp_path / 'tfrecords')
  signals = np.random.RandomState(0).rand(4, 16, 3).astype(np.float32)
  dataset_helper.write_dataset(data_dir, signals, np.zeros_like(signals), train_size=3)
  hp = argparse.Namespace(input_dir=data_dir, output_dir=str(tmp_path / 'runs'))
  dataset_helper.get_dataset_info(hp)
  assert hp.generated_dir == os.path.join(hp.output_dir, 'generated') and os.path.isdir(hp.generated_dir)


class _StubGan(object):
  """get_noise / generate of gan/algorithms/gan.py:29-30,92-97 on the host: sample i is filled with its running index"""

  def __init__(self, signal_shape):
    self.signal_shape, self.count, self.batches = tuple(signal_shape), 0, []

  def get_noise(self, batch_size):
    return np.zeros((batch_size, 4), np.float32)

  def generate(self, noise, denorm=False):
    assert denorm
    n = len(noise)
    self.batches.append(n)
    out = np.arange(self.count, self.count + n, dtype=np.float32).reshape((n,) + (1,) * len(self.signal_shape))
    self.count += n
    return np.broadcast_to(out, (n,) + self.signal_shape)


def test_generate_dataset_writes_generated_pkl(tmp_path):
  hp = argparse.Namespace(output_dir=str(tmp_path), signal_shape=(8, 3), verbose=0)
  gan = _StubGan(hp.signal_shape)
  filename = utils.generate_dataset(hp, gan, num_samples=250)
  assert filename == os.path.join(str(tmp_path), 'generated.pkl') and gan.batches == [100, 100, 50]
  with open(filename, 'rb') as file:
    content = pickle.load(file)
  assert list(content) == ['signals'] and content['signals'].dtype == np.float32 and content['signals'].shape == (250, 8, 3)
  np.testing.assert_array_equal(content['signals'][:, 0, 0], np.arange(250))


def test_array_format_helpers():
  hp = argparse.Namespace(sequence_length=16, num_neurons=5, validation_size=7)
  x = np.arange(7 * 16 * 5, dtype=np.float32).reshape(7, 16, 5)
  assert utils.get_array_format(x.shape, hp) == 'NWC' and utils.get_array_format((5, 16), hp) == 'CW'
  assert utils.set_array_format(x, 'NWC', hp) is x
  np.testing.assert_array_equal(utils.set_array_format(x, 'NCW', hp), x.transpose(0, 2, 1))
  np.testing.assert_array_equal(utils.set_array_format(x[0], 'CW', hp), x[0].T)      # what main.py:145-146 asks for
  torch = pytest.importorskip('torch')
  t = utils.set_array_format(torch.from_numpy(x), 'CWN', hp)
  assert tuple(t.shape) == (5, 16, 7) and float(t[2, 3, 4]) == x[4, 3, 2]
  with pytest.raises(AssertionError):
    utils.set_array_format(x, 'NW', hp)
  y = np.zeros((7, 5, 3))
  assert utils.swap_neuron_major(hp, y).shape == (5, 7, 3) and utils.swap_neuron_major(hp, np.zeros((5, 7, 3))).shape == (5, 7, 3)
  np.testing.assert_array_equal(utils.remove_nan(np.array([1.0, np.nan, 3.0])), [1.0, 3.0])


class _StubModel(object):

  def __init__(self, weights):
    self.weights = [w.copy() for w in weights]

  def get_weights(self):
    return [w.copy() for w in self.weights]

  def set_weights(self, weights):
    self.weights = [np.asarray(w).copy() for w in weights]


class _StubEngine(object):

  def __init__(self):
    self.state = {0: (np.full(3, 1.0, np.float32), np.full(3, 2.0, np.float32)),
                  1: (np.full(2, 3.0, np.float32), np.full(2, 4.0, np.float32))}

  def get_opt_state(self, which):
    m, v = self.state[which]
    return m.copy(), v.copy(), 0

  def set_opt_state(self, which, m, v, steps):
    self.state[which] = (np.asarray(m).copy(), np.asarray(v).copy())


def _stub_gan(seed):
  rng = np.random.RandomState(seed)
  return argparse.Namespace(generator=_StubModel([rng.rand(4, 3).astype(np.float32), rng.rand(3).astype(np.float32)]),
                            discriminator=_StubModel([rng.rand(2, 2).astype(np.float32)]),
                            gen_optimizer=argparse.Namespace(iterations=seed), dis_optimizer=argparse.Namespace(iterations=5 * seed),
                            engine=_StubEngine())


def test_checkpoint_layout_and_resume(tmp_path):
  """utils.py:116-152: output_dir/checkpoints/epoch-%03d.pkl with the reference's five keys; the newest epoch is resumed"""
  hp = argparse.Namespace(output_dir=str(tmp_path), verbose=0)
  a = _stub_gan(7)
  utils.save_models(hp, a, 9)
  a.gen_optimizer.iterations, a.dis_optimizer.iterations = 70, 350
  a.generator.weights[0] += 1.0
  a.engine.state[0] = (np.full(3, 9.0, np.float32), np.full(3, 8.0, np.float32))
  utils.save_models(hp, a, 10)
  assert sorted(os.listdir(hp.ckpt_dir)) == ['epoch-009.pkl', 'epoch-010.pkl'] and hp.ckpt_dir == os.path.join(str(tmp_path), 'checkpoints')
  with open(os.path.join(hp.ckpt_dir, 'epoch-010.pkl'), 'rb') as file:
    content = pickle.load(file)
  assert set(content) == {'epoch', 'gen_weights', 'dis_weights', 'gen_steps', 'dis_steps', 'b200_adam'}
  assert content['epoch'] == 10 and content['gen_steps'] == 70 and content['dis_steps'] == 350
  assert [w.dtype for w in content['gen_weights']] == [np.float32, np.float32]
  b = _stub_gan(1)
  hp2 = argparse.Namespace(output_dir=str(tmp_path), verbose=0)
  utils.load_models(hp2, b)
  assert hp2.start_epoch == 11 and b.gen_optimizer.iterations == 70 and b.dis_optimizer.iterations == 350
  for x, y in zip(b.generator.get_weights() + b.discriminator.get_weights(), a.generator.get_weights() + a.discriminator.get_weights()):
    np.testing.assert_array_equal(x, y)
  np.testing.assert_array_equal(b.engine.state[0][0], np.full(3, 9.0, np.float32))
  # a checkpoint written by the reference: no Adam moments, epoch numbers past the zero padding still order numerically
  with open(os.path.join(hp.ckpt_dir, 'epoch-1000.pkl'), 'wb') as file:
    pickle.dump({'epoch': 1000, 'gen_weights': a.generator.get_weights(), 'dis_weights': a.discriminator.get_weights(),
                 'gen_steps': 1, 'dis_steps': 2}, file)
  c = _stub_gan(2)
  before = c.engine.state[1][0].copy()
  hp3 = argparse.Namespace(output_dir=str(tmp_path), verbose=0)
  utils.load_models(hp3, c)
  assert hp3.start_epoch == 1001 and c.gen_optimizer.iterations == 1
  np.testing.assert_array_equal(c.engine.state[1][0], before)      # moments untouched
  # nothing to resume from
  hp4 = argparse.Namespace(output_dir=str(tmp_path / 'fresh'), verbose=0)
  utils.load_models(hp4, c)
  assert hp4.start_epoch == 0


def test_hparams_json_roundtrip(tmp_path):
  """utils.py:72-85"""
  import json
  hp = argparse.Namespace(output_dir=str(tmp_path), num_units=64, signal_shape=(2048, 102), lr=np.float32(1e-4), steps=np.int64(3))
  utils.save_hparams(hp)
  stored = json.load(open(os.path.join(str(tmp_path), 'hparams.json')))
  assert stored['signal_shape'] == [2048, 102] and stored['num_units'] == 64 and stored['steps'] == 3 and 'git_hash' in stored
  fresh = argparse.Namespace(output_dir=str(tmp_path), num_units=99)
  utils.load_hparams(fresh)
  assert fresh.num_units == 99 and fresh.signal_shape == [2048, 102] and abs(fresh.lr - 1e-4) < 1e-9
  assert abs(utils.denormalize(utils.normalize(3.0, -2.0, 6.0), -2.0, 6.0) - 3.0) < 1e-12
