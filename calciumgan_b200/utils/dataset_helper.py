"""TensorFlow-free reader (and writer) of the reference's dataset layout.

`dataset/generate_tfrecords.py:128-153,229-247` writes `train-XXX-of-YYY.record` / `validation-*.record` shards of
`tf.train.Example{signal: bytes(float32), spike: bytes(float32)}` plus `info.pkl`; `gan/utils/dataset_helper.py:113-206`
reads them back and fills the hparams fields the models use. This module does the same with the standard library:
TFRecord framing (u64 length, masked crc32c, payload, masked crc32c) and the three protobuf messages involved.
"""
import os
import pickle
import struct
from glob import glob

import numpy as np

# ------------------------------------------------------------------------------------------------ crc32c (Castagnoli)
_CRC_TABLE = None


def _crc_table():
  global _CRC_TABLE
  if _CRC_TABLE is None:
    tbl = []
    for i in range(256):
      c = i
      for _ in range(8):
        c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
      tbl.append(c)
    _CRC_TABLE = tbl
  return _CRC_TABLE


def crc32c(data):
  tbl, c = _crc_table(), 0xFFFFFFFF
  for b in data:
    c = tbl[(c ^ b) & 0xFF] ^ (c >> 8)
  return c ^ 0xFFFFFFFF


def masked_crc32c(data):
  c = crc32c(data)
  return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ protobuf wire format
def _read_varint(buf, pos):
  result, shift = 0, 0
  while True:
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7


def _fields(buf):
  """Yield (field_number, wire_type, value) of one message; length-delimited values are memoryview slices."""
  pos, n = 0, len(buf)
  while pos < n:
    key, pos = _read_varint(buf, pos)
    num, wt = key >> 3, key & 7
    if wt == 0:
      val, pos = _read_varint(buf, pos)
    elif wt == 2:
      ln, pos = _read_varint(buf, pos)
      val = buf[pos:pos + ln]
      pos += ln
    elif wt == 1:
      val = buf[pos:pos + 8]
      pos += 8
    elif wt == 5:
      val = buf[pos:pos + 4]
      pos += 4
    else:
      raise ValueError('unsupported protobuf wire type %d' % wt)
    yield num, wt, val


def parse_example(record):
  """tf.train.Example -> {feature name: first bytes_list value}. Example.features = 1; Features.feature = 1 (map
  entry: key = 1, value = 2); Feature.bytes_list = 1; BytesList.value = 1."""
  out = {}
  buf = memoryview(record)
  for num, wt, features in _fields(buf):
    if num != 1 or wt != 2:
      continue
    for num2, wt2, entry in _fields(features):
      if num2 != 1 or wt2 != 2:
        continue
      key, value = None, None
      for num3, wt3, v in _fields(entry):
        if num3 == 1 and wt3 == 2:
          key = bytes(v).decode('utf-8')
        elif num3 == 2 and wt3 == 2:
          for num4, wt4, blist in _fields(v):
            if num4 == 1 and wt4 == 2:          # bytes_list
              for num5, wt5, val in _fields(blist):
                if num5 == 1 and wt5 == 2:
                  value = val
                  break
      if key is not None and value is not None:
        out[key] = value
  return out


def _varint(n):
  out = bytearray()
  while True:
    b = n & 0x7F
    n >>= 7
    out.append(b | (0x80 if n else 0))
    if not n:
      return bytes(out)


def _ld(field, payload):
  return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def serialize_example(signal, spike):
  """generate_tfrecords.py:132-139 without TensorFlow."""
  entries = b''
  for key, arr in (('signal', signal), ('spike', spike)):
    feature = _ld(1, _ld(1, np.ascontiguousarray(arr, np.float32).tobytes()))
    entries += _ld(1, _ld(1, key.encode()) + _ld(2, feature))
  return _ld(1, entries)


# ------------------------------------------------------------------------------------------------ TFRecord framing
def read_records(path, verify_crc=False):
  with open(path, 'rb') as f:
    while True:
      header = f.read(12)
      if len(header) < 12:
        return
      (length,), (len_crc,) = struct.unpack('<Q', header[:8]), struct.unpack('<I', header[8:])
      if verify_crc and masked_crc32c(header[:8]) != len_crc:
        raise IOError('corrupted TFRecord length in %s' % path)
      data = f.read(length)
      (data_crc,) = struct.unpack('<I', f.read(4))
      if verify_crc and masked_crc32c(data) != data_crc:
        raise IOError('corrupted TFRecord payload in %s' % path)
      yield data


def write_records(path, payloads):
  with open(path, 'wb') as f:
    for data in payloads:
      header = struct.pack('<Q', len(data))
      f.write(header + struct.pack('<I', masked_crc32c(header)) + data + struct.pack('<I', masked_crc32c(data)))


# ------------------------------------------------------------------------------------------------ dataset
def get_dataset_info(hparams):
  """gan/utils/dataset_helper.py:113-144."""
  with open(os.path.join(hparams.input_dir, 'info.pkl'), 'rb') as file:
    info = pickle.load(file)
  hparams.train_files = os.path.join(hparams.input_dir, 'train-*.record')
  hparams.validation_files = os.path.join(hparams.input_dir, 'validation-*.record')
  for key in ('train_size', 'validation_size', 'signal_shape', 'spike_shape', 'sequence_length', 'num_neurons',
              'num_channels', 'num_train_shards', 'num_validation_shards', 'buffer_size', 'normalize', 'fft', 'conv2d'):
    setattr(hparams, key, info[key])
  hparams.signal_shape = tuple(hparams.signal_shape)
  if hparams.normalize:
    hparams.signals_min = float(info['signals_min'])
    hparams.signals_max = float(info['signals_max'])
  set_generated_dir(hparams)
  return info


def set_generated_dir(hparams):
  """gan/utils/dataset_helper.py:139-141: generated signals go to output_dir/generated"""
  hparams.generated_dir = os.path.join(hparams.output_dir, 'generated')
  os.makedirs(hparams.generated_dir, exist_ok=True)
  hparams.validation_cache = os.path.join(hparams.generated_dir, 'validation.h5')


def _load_split(pattern, signal_shape, spike_shape=None):
  """every record of the shards matching `pattern`, in file order: signals (N,) + signal_shape and, when `spike_shape`
  is given, spikes (N,) + spike_shape (dataset_helper.py:157-164)"""
  signals, spikes = [], []
  for path in sorted(glob(pattern)):
    for record in read_records(path):
      ex = parse_example(record)
      signals.append(np.frombuffer(ex['signal'], dtype=np.float32).reshape(signal_shape))
      if spike_shape is not None:
        spikes.append(np.frombuffer(ex['spike'], dtype=np.float32).reshape(spike_shape))
  if not signals:
    raise IOError('no records match %s' % pattern)
  return np.stack(signals), (np.stack(spikes) if spike_shape is not None else None)


def shuffle_buffer_order(n, buffer_size, rng):
  """Order in which `tf.data.Dataset.shuffle(buffer_size)` (dataset_helper.py:172) emits n elements: a buffer of
  `buffer_size` elements is filled in input order, a uniformly chosen slot is emitted and refilled with the next input
  element. With buffer_size >= n this is a uniform permutation; with a smaller buffer element i cannot appear before
  output position i - buffer_size + 1 (the locality a windowed shuffle has). Same algorithm, not the same random stream."""
  buffer_size = max(1, int(buffer_size))
  buf = list(range(min(buffer_size, n)))
  nxt = len(buf)
  order = np.empty(n, dtype=np.int64)
  for k in range(n):
    j = int(rng.randint(len(buf)))
    order[k] = buf[j]
    if nxt < n:
      buf[j] = nxt
      nxt += 1
    else:
      buf[j] = buf[-1]
      buf.pop()
  return order


def shard_for_rank(n, rank, world_size):
  """Indices of the samples rank `rank` trains on under batch-sharded data parallelism: a strided slice of the first
  floor(n / world) * world samples, so every rank has the same number of samples -- hence the same number of batches and
  the same (possibly ragged) last batch size, which the gradient exchange in every `gan.train` call requires."""
  per = n // world_size
  if per == 0:
    raise ValueError('%d training samples cannot be split over %d ranks' % (n, world_size))
  return np.arange(per * world_size)[rank::world_size]


class _Batches(object):
  """Re-iterable batch source: shuffle(buffer_size) (train) + batch without drop_remainder (dataset_helper.py:171-181).
  Yields (signals, spikes or None) like the reference's (signal, spike) pairs."""

  def __init__(self, signals, batch_size, shuffle, seed=1234, spikes=None, buffer_size=None):
    self.signals, self.spikes, self.batch_size, self.shuffle = signals, spikes, batch_size, shuffle
    self.buffer_size = len(signals) if buffer_size is None else buffer_size
    self.rng = np.random.RandomState(seed)

  def __len__(self):
    return int(np.ceil(len(self.signals) / self.batch_size))

  def __iter__(self):
    n = len(self.signals)
    idx = shuffle_buffer_order(n, self.buffer_size, self.rng) if self.shuffle else np.arange(n)
    for i in range(0, n, self.batch_size):
      sel = idx[i:i + self.batch_size]
      yield self.signals[sel], (None if self.spikes is None else self.spikes[sel])


def cache_validation_set(hparams, validation_ds):
  """dataset_helper.py:12-31: the de-normalised validation signals (float32) and spikes (int8) go to
  generated_dir/validation.h5 once, next to the generated signals they are compared with."""
  from . import h5_helper, utils
  if h5_helper.exists(hparams.validation_cache):
    return
  for signal, spike in validation_ds:
    signal = utils.reverse_preprocessing(hparams, np.asarray(signal))
    h5_helper.write(hparams.validation_cache, {'signals': signal.astype(np.float32), 'spikes': np.asarray(spike).astype(np.int8)})


def get_surrogate_dataset(hparams):
  """dataset_helper.py:53-110: input_dir/training.pkl {'signals': (trials, neurons, time), 'spikes'} -> NWC signals scaled
  to [0, 1] by the global min / max, the first 8192 trials train (shuffle buffer 2048), the rest validate."""
  filename = os.path.join(hparams.input_dir, 'training.pkl')
  if not os.path.exists(filename):
    print('training dataset {} not found'.format(filename))
    exit()
  with open(filename, 'rb') as file:
    data = pickle.load(file)
  signals = np.transpose(np.asarray(data['signals'], np.float32), axes=[0, 2, 1])
  spikes = np.asarray(data['spikes'])
  hparams.signals_min, hparams.signals_max = float(np.min(signals)), float(np.max(signals))
  signals = (signals - hparams.signals_min) / (hparams.signals_max - hparams.signals_min)
  train_size = 8192
  hparams.train_size, hparams.validation_size = len(signals[:train_size]), len(signals[train_size:])
  hparams.signal_shape = tuple(signals.shape[1:])
  hparams.spike_shape = tuple(spikes.shape[1:])
  hparams.sequence_length, hparams.num_neurons, hparams.num_channels = signals.shape[1], signals.shape[-1], signals.shape[-1]
  hparams.normalize, hparams.fft, hparams.conv2d = True, False, False
  set_generated_dir(hparams)
  return (_Batches(signals[:train_size], hparams.batch_size, True, spikes=spikes[:train_size], buffer_size=2048),
          _Batches(signals[train_size:], hparams.batch_size, False, spikes=spikes[train_size:]))


def get_dataset(hparams, summary=None):
  """gan/utils/dataset_helper.py:185-206; returns re-iterable (train_ds, validation_ds) of (signal, spike) batches.
  Under data parallelism (world_size > 1) the training split is sharded by `shard_for_rank`; validation is not."""
  hparams.noise_shape = (hparams.noise_dim,)
  rank, world = int(getattr(hparams, 'rank', 0)), int(getattr(hparams, 'world_size', 1))
  if getattr(hparams, 'surrogate_ds', False):
    train_ds, validation_ds = get_surrogate_dataset(hparams)
  else:
    if not os.path.exists(hparams.input_dir):
      print('input directory {} cannot be found'.format(hparams.input_dir))
      exit()
    get_dataset_info(hparams)
    train, train_spikes = _load_split(hparams.train_files, hparams.signal_shape, hparams.spike_shape)      # == ds.cache()
    val, val_spikes = _load_split(hparams.validation_files, hparams.signal_shape, hparams.spike_shape)
    train_ds = _Batches(train, hparams.batch_size, True, spikes=train_spikes, buffer_size=hparams.buffer_size)
    validation_ds = _Batches(val, hparams.batch_size, False, spikes=val_spikes)
    if getattr(hparams, 'save_generated', '') and rank == 0:
      cache_validation_set(hparams, validation_ds)
  if world > 1:
    sel = shard_for_rank(len(train_ds.signals), rank, world)
    train_ds = _Batches(train_ds.signals[sel], hparams.batch_size, True, seed=1234 + rank,
                        spikes=None if train_ds.spikes is None else train_ds.spikes[sel], buffer_size=train_ds.buffer_size)
    hparams.train_size = len(sel)
  hparams.train_steps = int(np.ceil(hparams.train_size / hparams.batch_size))
  hparams.validation_steps = int(np.ceil(hparams.validation_size / hparams.batch_size))
  return train_ds, validation_ds


def write_dataset(output_dir, signals, spikes, train_size, num_per_shard=1100, normalize=True):
  """Minimal stand-in for dataset/generate_tfrecords.py:186-252 (same files, same info.pkl keys)."""
  os.makedirs(output_dir, exist_ok=True)
  signals = np.asarray(signals, np.float32)
  spikes = np.asarray(spikes, np.float32)
  smin, smax = float(signals.min()), float(signals.max())
  if normalize:
    signals = (signals - smin) / (smax - smin)
  splits = {'train': np.arange(train_size), 'validation': np.arange(train_size, len(signals))}
  shards = {}
  for mode, idx in splits.items():
    n = max(1, int(np.ceil(len(idx) / num_per_shard)))
    shards[mode] = n
    for s, part in enumerate(np.array_split(idx, n)):
      path = os.path.join(output_dir, '{}-{:03d}-of-{:03d}.record'.format(mode, s + 1, n))
      write_records(path, (serialize_example(signals[i], spikes[i]) for i in part))
  info = {
      'train_size': train_size, 'validation_size': len(signals) - train_size,
      'signal_shape': signals.shape[1:], 'spike_shape': spikes.shape[1:], 'sequence_length': signals.shape[1],
      'num_neurons': signals.shape[-1], 'num_channels': signals.shape[-1], 'num_train_shards': shards['train'],
      'num_validation_shards': shards['validation'], 'buffer_size': min(num_per_shard, train_size),
      'normalize': normalize, 'stride': 2, 'fft': False, 'conv2d': False,
  }
  if normalize:
    info['signals_min'], info['signals_max'] = smin, smax
  with open(os.path.join(output_dir, 'info.pkl'), 'wb') as file:
    pickle.dump(info, file)
  return info
