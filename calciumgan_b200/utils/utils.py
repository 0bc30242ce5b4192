"""Host utilities under the names the reference's callers use (gan/utils/utils.py), written against the same on-disk
formats:

  checkpoints  output_dir/checkpoints/epoch-%03d.pkl = {'epoch', 'gen_weights', 'dis_weights', 'gen_steps', 'dis_steps'}
               with fp32 numpy arrays in Keras get_weights() order (utils.py:116-152). The reference pickles tf.Variable
               step counters; plain ints are written here (its loader assigns whatever it finds). Adam moments, which
               the reference does not save, travel under the extra key 'b200_adam' that its loader ignores.
  hparams      output_dir/hparams.json (utils.py:72-85)
  generated    generated_dir/epoch%03d_signals.h5 + generated_dir/info.pkl (utils.py:93-113), output_dir/generated.pkl
               (utils.py:191-207)
"""
import json
import os
import pickle
import re
import subprocess
from glob import glob

import numpy as np

from .. import _lib as L

ADAM_KEY = 'b200_adam'


# ------------------------------------------------------------------------------------------------ signal scaling
def normalize(x, x_min, x_max):
  """[x_min, x_max] -> [0, 1] (utils.py:25-27)"""
  span = x_max - x_min
  return (x - x_min) / span


def denormalize(x, x_min, x_max):
  """[0, 1] -> [x_min, x_max], the inverse of `normalize` (utils.py:30-32)"""
  span = x_max - x_min
  return x * span + x_min


def reverse_preprocessing(hparams, x):
  """Undo the dataset preprocessing so that generated signals are on the scale of the recordings (utils.py:50-63). Only
  the 1-D, non-FFT layout is in scope (the conv2d / fft variants are not built), so this is the de-normalisation."""
  for variant in ('conv2d', 'fft'):
    if getattr(hparams, variant, False):
      raise NotImplementedError('%s datasets are out of scope of calciumgan_b200' % variant)
  return denormalize(x, hparams.signals_min, hparams.signals_max) if hparams.normalize else x


# ------------------------------------------------------------------------------------------------ generated signals
def _to_host(x):
  return x.detach().float().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)


def _note_epoch(index_path, epoch, entry):
  """generated_dir/info.pkl maps epoch -> entry; the first entry of an epoch stays"""
  index = {}
  if os.path.exists(index_path):
    with open(index_path, 'rb') as f:
      index = pickle.load(f)
  if epoch in index:
    return
  index[epoch] = entry
  with open(index_path, 'wb') as f:
    pickle.dump(index, f)


def save_fake_signals(hparams, epoch, signals):
  """Append one validation batch of generated signals -- on the recordings' scale, float32, NWC -- to
  `generated_dir/epoch%03d_signals.h5` (key 'signals') and note {'global_step', 'filename'} for the epoch in
  `generated_dir/info.pkl` (utils.py:93-113). `signals` may be the device tensor `gan.validate` returns."""
  from . import h5_helper
  target = os.path.join(hparams.generated_dir, 'epoch{:03d}_signals.h5'.format(epoch))
  batch = reverse_preprocessing(hparams, _to_host(signals)).astype(np.float32)
  h5_helper.write(target, {'signals': batch})
  _note_epoch(os.path.join(hparams.generated_dir, 'info.pkl'), epoch,
              {'global_step': hparams.global_step, 'filename': target})


def save_generated_at(hparams, epoch):
  """The epochs after which main.py saves generated signals (main.py:81-84 of the reference): with --save_generated all
  every 10th epoch and the last one, with --save_generated last only the last one."""
  is_last = epoch == hparams.epochs - 1
  if hparams.save_generated == 'all':
    return is_last or epoch % 10 == 0
  return hparams.save_generated == 'last' and is_last


def generate_dataset(hparams, gan, num_samples=1000, batch_size=100):
  """`num_samples` generator outputs on the recordings' scale, drawn `batch_size` at a time, pickled as
  {'signals': float32 (num_samples,) + signal_shape} to output_dir/generated.pkl (utils.py:191-207; main.py:219-221 asks
  for 2e6 samples on surrogate datasets). The last batch is cut to size (the reference needs num_samples % 100 == 0)."""
  out = np.zeros((num_samples,) + tuple(hparams.signal_shape), dtype=np.float32)
  done = 0
  while done < num_samples:
    n = min(batch_size, num_samples - done)
    out[done:done + n] = _to_host(gan.generate(gan.get_noise(n), denorm=True))
    done += n
  target = os.path.join(hparams.output_dir, 'generated.pkl')
  with open(target, 'wb') as f:
    pickle.dump({'signals': out}, f, protocol=4)      # protocol 4: arrays beyond 4 GiB
  if getattr(hparams, 'verbose', 0):
    print('save {} samples to {}'.format(num_samples, target))
  return target


# ------------------------------------------------------------------------------------------------ array layouts
def swap_neuron_major(hparams, array):
  """(validation_size, num_neurons, ...) -> (num_neurons, validation_size, ...); any other layout passes through
  (utils.py:86-89)"""
  trial_major = tuple(array.shape[:2]) == (hparams.validation_size, hparams.num_neurons)
  return np.swapaxes(array, 0, 1) if trial_major else array


def get_array_format(shape, hparams):
  """One letter per axis: W where the extent is the sequence length, C where it is the number of neurons, N otherwise
  (utils.py:154-165)"""
  assert len(shape) <= 3
  letter = {hparams.num_neurons: 'C', hparams.sequence_length: 'W'}     # W wins when both extents coincide
  return ''.join(letter.get(extent, 'N') for extent in shape)


def set_array_format(array, data_format, hparams):
  """Permute a numpy array or torch tensor into `data_format`, e.g. one (W, C) block of traces to 'CW' (utils.py:168-184)"""
  assert len(array.shape) == len(data_format)
  have = get_array_format(array.shape, hparams)
  assert set(have) == set(data_format)
  if have == data_format:
    return array
  axes = [have.index(letter) for letter in data_format]
  return array.permute(*axes) if hasattr(array, 'permute') else np.transpose(array, axes=axes)


def remove_nan(array):
  """the entries of `array` that are not NaN, flattened (utils.py:187-188)"""
  return array[~np.isnan(array)]


# ------------------------------------------------------------------------------------------------ hparams.json
def get_current_git_hash():
  """`git describe --always` (utils.py:66-69); 'unknown' outside a git checkout instead of raising"""
  try:
    out = subprocess.check_output(['git', 'describe', '--always'], stderr=subprocess.DEVNULL)
  except Exception:
    return 'unknown'
  return out.strip().decode()


def _json_plain(value):
  if isinstance(value, np.integer):
    return int(value)
  if isinstance(value, np.floating):
    return float(value)
  if isinstance(value, np.ndarray):
    return value.tolist()
  raise TypeError('hparams field of type %s is not JSON serialisable' % type(value).__name__)


def _hparams_path(hparams):
  return os.path.join(hparams.output_dir, 'hparams.json')


def save_hparams(hparams):
  """Every field of the Namespace plus `git_hash` -> output_dir/hparams.json (utils.py:72-75). Tuples become lists, as
  with the reference's json.dump; numpy scalars and arrays are converted."""
  hparams.git_hash = get_current_git_hash()
  with open(_hparams_path(hparams), 'w') as f:
    json.dump(vars(hparams), f, default=_json_plain)


def load_hparams(hparams):
  """Fill in, from output_dir/hparams.json, the fields the Namespace does not have yet; fields it has win (utils.py:78-84)"""
  with open(_hparams_path(hparams), 'r') as f:
    stored = json.load(f)
  for key in stored.keys() - vars(hparams).keys():
    setattr(hparams, key, stored[key])


# ------------------------------------------------------------------------------------------------ checkpoints
def _checkpoint_dir(hparams):
  if not hasattr(hparams, 'ckpt_dir'):
    hparams.ckpt_dir = os.path.join(hparams.output_dir, 'checkpoints')
  return hparams.ckpt_dir


def _latest_checkpoint(directory):
  """the epoch-* file of the highest epoch (the reference takes the lexicographically last name, which is the same file
  for its zero-padded names)"""
  best = None
  for path in glob(os.path.join(directory, 'epoch-*')):
    m = re.search(r'epoch-(\d+)', os.path.basename(path))
    key = (int(m.group(1)) if m else -1, path)
    if best is None or key > best:
      best = key
  return None if best is None else best[1]


def save_models(hparams, gan, epoch, save_optimizer_state=True):
  """utils.py:116-133"""
  directory = _checkpoint_dir(hparams)
  os.makedirs(directory, exist_ok=True)
  state = dict(epoch=epoch,
               gen_weights=gan.generator.get_weights(), dis_weights=gan.discriminator.get_weights(),
               gen_steps=int(gan.gen_optimizer.iterations), dis_steps=int(gan.dis_optimizer.iterations))
  if save_optimizer_state:
    moments = {}
    for tag, which in (('gen', L.GENERATOR), ('dis', L.DISCRIMINATOR)):
      moments[tag + '_m'], moments[tag + '_v'], _ = gan.engine.get_opt_state(which)
    state[ADAM_KEY] = moments
  target = os.path.join(directory, 'epoch-{:03d}.pkl'.format(epoch))
  with open(target, 'wb') as f:
    pickle.dump(state, f)
  if getattr(hparams, 'verbose', 0):
    print('Saved checkpoint to {}'.format(target))


def load_models(hparams, gan):
  """Resume from the newest checkpoint, if any: weights, step counters and `hparams.start_epoch` (utils.py:136-152); Adam
  moments when the checkpoint was written by this package (the reference restarts them from zero, utils.py:126-127,148-149)."""
  hparams.start_epoch = 0
  source = _latest_checkpoint(_checkpoint_dir(hparams))
  if source is None:
    return
  with open(source, 'rb') as f:
    state = pickle.load(f)
  hparams.start_epoch = state['epoch'] + 1
  for model, key in ((gan.generator, 'gen_weights'), (gan.discriminator, 'dis_weights')):
    model.set_weights(state[key])
  moments = state.get(ADAM_KEY)
  if moments is not None:
    for tag, which in (('gen', L.GENERATOR), ('dis', L.DISCRIMINATOR)):
      gan.engine.set_opt_state(which, moments[tag + '_m'], moments[tag + '_v'], 0)
  gan.gen_optimizer.iterations = state['gen_steps']
  gan.dis_optimizer.iterations = state['dis_steps']
  if getattr(hparams, 'verbose', 0):
    print('\n\nRestored checkpoint at {}\n\n'.format(source))
