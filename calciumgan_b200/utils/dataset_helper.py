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


def _load_split(pattern, signal_shape):
  signals = []
  for path in sorted(glob(pattern)):
    for record in read_records(path):
      ex = parse_example(record)
      signals.append(np.frombuffer(ex['signal'], dtype=np.float32).reshape(signal_shape))
  if not signals:
    raise IOError('no records match %s' % pattern)
  return np.stack(signals)


class _Batches(object):
  """Re-iterable batch source: shuffle (train) + batch without drop_remainder (dataset_helper.py:171-181)."""

  def __init__(self, signals, batch_size, shuffle, seed=1234):
    self.signals, self.batch_size, self.shuffle = signals, batch_size, shuffle
    self.rng = np.random.RandomState(seed)

  def __len__(self):
    return int(np.ceil(len(self.signals) / self.batch_size))

  def __iter__(self):
    idx = self.rng.permutation(len(self.signals)) if self.shuffle else np.arange(len(self.signals))
    for i in range(0, len(idx), self.batch_size):
      yield self.signals[idx[i:i + self.batch_size]], None


def get_dataset(hparams, summary=None):
  """gan/utils/dataset_helper.py:185-206 for the TFRecord layout; returns re-iterable (train_ds, validation_ds)."""
  hparams.noise_shape = (hparams.noise_dim,)
  get_dataset_info(hparams)
  train = _load_split(hparams.train_files, hparams.signal_shape)          # == ds.cache()
  val = _load_split(hparams.validation_files, hparams.signal_shape)
  hparams.train_steps = int(np.ceil(hparams.train_size / hparams.batch_size))
  hparams.validation_steps = int(np.ceil(hparams.validation_size / hparams.batch_size))
  return _Batches(train, hparams.batch_size, True), _Batches(val, hparams.batch_size, False)


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
