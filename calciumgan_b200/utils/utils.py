"""Checkpoint + small helpers with the reference's names and on-disk layout
(gan/utils/utils.py:30-32,116-152): `output_dir/checkpoints/epoch-%03d.pkl` holding
{'epoch', 'gen_weights', 'dis_weights', 'gen_steps', 'dis_steps'} with fp32 numpy arrays in
Keras get_weights() order. The reference pickles tf.Variable step counters; plain ints are
written here (its loader assigns whatever it finds). Adam moments, which the reference does
not save, go under extra keys the reference's loader ignores."""
import os
import pickle
from glob import glob

import numpy as np

from .. import _lib as L


def normalize(x, x_min, x_max):
  ''' scale x to be between 0 and 1 '''
  return (x - x_min) / (x_max - x_min)


def denormalize(x, x_min, x_max):
  ''' re-scale signals back to its original range '''
  return x * (x_max - x_min) + x_min


def reverse_preprocessing(hparams, x):
  """gan/utils/utils.py:50-63: undo the dataset preprocessing so that generated signals match the raw recordings. Only the
  1-D, non-FFT layout is in scope (conv2d / fft variants are not built)."""
  if getattr(hparams, 'conv2d', False) or getattr(hparams, 'fft', False):
    raise NotImplementedError('conv2d / fft datasets are out of scope of calciumgan_b200')
  if hparams.normalize:
    x = denormalize(x, x_min=hparams.signals_min, x_max=hparams.signals_max)
  return x


def save_fake_signals(hparams, epoch, signals):
  """gan/utils/utils.py:93-113: append one validation batch of generated signals (de-normalised, float32, NWC) to
  `generated_dir/epoch%03d_signals.h5` under the key 'signals', and record {epoch: {'global_step', 'filename'}} in
  `generated_dir/info.pkl` the first time the epoch is seen. `signals` may be a device tensor (what `gan.validate` returns)."""
  from . import h5_helper
  if hasattr(signals, 'detach'):
    signals = signals.detach().float().cpu().numpy()
  signals = reverse_preprocessing(hparams, np.asarray(signals))
  filename = os.path.join(hparams.generated_dir, 'epoch{:03d}_signals.h5'.format(epoch))
  h5_helper.write(filename, {'signals': signals.astype(np.float32)})
  info_filename = os.path.join(hparams.generated_dir, 'info.pkl')
  info = {}
  if os.path.exists(info_filename):
    with open(info_filename, 'rb') as file:
      info = pickle.load(file)
  if epoch not in info:
    info[epoch] = {'global_step': hparams.global_step, 'filename': filename}
    with open(info_filename, 'wb') as file:
      pickle.dump(info, file)


def save_generated_at(hparams, epoch):
  """main.py:81-84 of the reference: 'all' -> every 10th epoch and the last one, 'last' -> the last epoch only"""
  last = epoch == hparams.epochs - 1
  return (hparams.save_generated == 'all' and (epoch % 10 == 0 or last)) or (hparams.save_generated == 'last' and last)


def swap_neuron_major(hparams, array):
  """gan/utils/utils.py:86-89: (validation_size, num_neurons, ...) -> neuron-major; anything else is returned as is"""
  if tuple(array.shape[:2]) == (hparams.validation_size, hparams.num_neurons):
    return np.swapaxes(array, 0, 1)
  return array


def get_array_format(shape, hparams):
  """gan/utils/utils.py:154-165: one letter per axis -- W = sequence length, C = number of neurons, N = anything else"""
  assert len(shape) <= 3
  return ''.join('W' if s == hparams.sequence_length else 'C' if s == hparams.num_neurons else 'N' for s in shape)


def set_array_format(array, data_format, hparams):
  """gan/utils/utils.py:168-184: permute `array` (numpy or torch) to `data_format`, e.g. a (W, C) trace block to 'CW'"""
  assert len(array.shape) == len(data_format)
  current = get_array_format(array.shape, hparams)
  assert set(current) == set(data_format)
  if data_format == current:
    return array
  perm = [current.index(s) for s in data_format]
  return array.permute(*perm) if hasattr(array, 'permute') else np.transpose(array, axes=perm)


def remove_nan(array):
  """gan/utils/utils.py:187-188"""
  return array[np.logical_not(np.isnan(array))]


def generate_dataset(hparams, gan, num_samples=1000, batch_size=100):
  """gan/utils/utils.py:191-207: `num_samples` de-normalised generator outputs in batches of 100 ->
  output_dir/generated.pkl {'signals': float32 (num_samples,) + signal_shape} (main.py:219-221 calls it with 2e6 samples
  for surrogate datasets). The last batch is cut to size (the reference requires num_samples % 100 == 0)."""
  generated = np.zeros((num_samples,) + tuple(hparams.signal_shape), dtype=np.float32)
  for i in range(0, num_samples, batch_size):
    n = min(batch_size, num_samples - i)
    signals = gan.generate(gan.get_noise(n), denorm=True)
    if hasattr(signals, 'detach'):
      signals = signals.detach().float().cpu().numpy()
    generated[i:i + n] = signals
  filename = os.path.join(hparams.output_dir, 'generated.pkl')
  with open(filename, 'wb') as file:
    pickle.dump({'signals': generated}, file, protocol=4)      # protocol 4: arrays beyond 4 GiB
  if getattr(hparams, 'verbose', 0):
    print('save {} samples to {}'.format(num_samples, filename))
  return filename


def get_current_git_hash():
  """gan/utils/utils.py:66-69; 'unknown' outside a git checkout instead of raising."""
  import subprocess
  try:
    return subprocess.check_output(['git', 'describe', '--always'], stderr=subprocess.DEVNULL).strip().decode()
  except Exception:
    return 'unknown'


def save_hparams(hparams):
  """gan/utils/utils.py:72-75: output_dir/hparams.json with every hparams field (+ git_hash). Tuples become lists, as
  with the reference's json.dump; numpy scalars are converted."""
  import json
  hparams.git_hash = get_current_git_hash()

  def plain(v):
    if isinstance(v, (np.integer,)):
      return int(v)
    if isinstance(v, (np.floating,)):
      return float(v)
    if isinstance(v, np.ndarray):
      return v.tolist()
    raise TypeError('hparams field of type %s is not JSON serialisable' % type(v).__name__)

  with open(os.path.join(hparams.output_dir, 'hparams.json'), 'w') as file:
    json.dump(hparams.__dict__, file, default=plain)


def load_hparams(hparams):
  """gan/utils/utils.py:78-84: fill in the fields the Namespace does not have yet."""
  import json
  filename = os.path.join(hparams.output_dir, 'hparams.json')
  with open(filename, 'r') as file:
    content = json.load(file)
  for key, value in content.items():
    if not hasattr(hparams, key):
      setattr(hparams, key, value)


def save_models(hparams, gan, epoch, save_optimizer_state=True):
  if not hasattr(hparams, 'ckpt_dir'):
    hparams.ckpt_dir = os.path.join(hparams.output_dir, 'checkpoints')
  if not os.path.exists(hparams.ckpt_dir):
    os.makedirs(hparams.ckpt_dir)
  filename = os.path.join(hparams.ckpt_dir, 'epoch-{:03d}.pkl'.format(epoch))

  with open(filename, 'wb') as file:
    content = {
        'epoch': epoch,
        'gen_weights': gan.generator.get_weights(),
        'dis_weights': gan.discriminator.get_weights(),
        'gen_steps': int(gan.gen_optimizer.iterations),
        'dis_steps': int(gan.dis_optimizer.iterations)
    }
    if save_optimizer_state:
      gm, gv, _ = gan.engine.get_opt_state(L.GENERATOR)
      dm, dv, _ = gan.engine.get_opt_state(L.DISCRIMINATOR)
      content['b200_adam'] = {'gen_m': gm, 'gen_v': gv, 'dis_m': dm, 'dis_v': dv}
    pickle.dump(content, file)

  if getattr(hparams, 'verbose', 0):
    print('Saved checkpoint to {}'.format(filename))


def load_models(hparams, gan):
  if not hasattr(hparams, 'ckpt_dir'):
    hparams.ckpt_dir = os.path.join(hparams.output_dir, 'checkpoints')

  hparams.start_epoch = 0
  filenames = glob(os.path.join(hparams.ckpt_dir, 'epoch-*'))
  if filenames:
    filename = sorted(filenames)[-1]
    with open(filename, 'rb') as file:
      ckpt = pickle.load(file)
    hparams.start_epoch = ckpt['epoch'] + 1
    gan.generator.set_weights(ckpt['gen_weights'])
    gan.discriminator.set_weights(ckpt['dis_weights'])
    adam = ckpt.get('b200_adam')
    if adam is not None:   # the reference restarts the moments from zero (utils.py:126-127,148-149)
      gan.engine.set_opt_state(L.GENERATOR, adam['gen_m'], adam['gen_v'], 0)
      gan.engine.set_opt_state(L.DISCRIMINATOR, adam['dis_m'], adam['dis_v'], 0)
    gan.gen_optimizer.iterations = ckpt['gen_steps']
    gan.dis_optimizer.iterations = ckpt['dis_steps']

    if getattr(hparams, 'verbose', 0):
      print('\n\nRestored checkpoint at {}\n\n'.format(filename))
