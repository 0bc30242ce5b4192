"""TF-free TFRecord reader/writer for the reference's dataset layout (dataset/generate_tfrecords.py,
gan/utils/dataset_helper.py): framing, protobuf payload, info.pkl -> hparams, batching without drop_remainder."""
import argparse
import os
import pickle
import struct

import numpy as np
import pytest

from calciumgan_b200.utils import dataset_helper as D


def test_crc32c_known_answers():
  assert D.crc32c(b'') == 0
  assert D.crc32c(b'123456789') == 0xE3069283          # standard CRC-32C check value
  assert D.crc32c(bytes(32)) == 0x8A9136AA               # RFC 3720 B.4: 32 bytes of zeros


def test_example_roundtrip_and_wire_format():
  rng = np.random.RandomState(0)
  sig, spk = rng.rand(8, 3).astype(np.float32), rng.rand(8, 3).astype(np.float32)
  rec = D.serialize_example(sig, spk)
  # tf.train.Example: field 1 (features), length-delimited
  assert rec[0] == 0x0A
  ex = D.parse_example(rec)
  assert set(ex) == {'signal', 'spike'}
  np.testing.assert_array_equal(np.frombuffer(ex['signal'], np.float32).reshape(8, 3), sig)
  np.testing.assert_array_equal(np.frombuffer(ex['spike'], np.float32).reshape(8, 3), spk)


def test_tfrecord_framing(tmp_path):
  path = str(tmp_path / 'x.record')
  payloads = [b'abc', b'', bytes(range(200))]
  D.write_records(path, payloads)
  assert list(D.read_records(path, verify_crc=True)) == payloads
  raw = open(path, 'rb').read()
  assert struct.unpack('<Q', raw[:8])[0] == 3 and len(raw) == sum(16 + len(p) for p in payloads)


def test_dataset_layout_to_hparams(tmp_path):
  rng = np.random.RandomState(1)
  signals = rng.rand(11, 64, 6).astype(np.float32) * 3 - 1
  spikes = (rng.rand(11, 64, 6) > 0.9).astype(np.float32)
  D.write_dataset(str(tmp_path), signals, spikes, train_size=8, num_per_shard=3)
  hp = argparse.Namespace(input_dir=str(tmp_path), output_dir=str(tmp_path / 'runs'), batch_size=3, noise_dim=4)
  train_ds, val_ds = D.get_dataset(hp)
  assert hp.signal_shape == (64, 6) and hp.num_channels == 6 and hp.normalize and hp.noise_shape == (4,)
  assert hp.train_size == 8 and hp.validation_size == 3 and hp.train_steps == 3 and hp.num_train_shards == 3
  assert abs(hp.signals_min - float(signals.min())) < 1e-6 and abs(hp.signals_max - float(signals.max())) < 1e-6
  sizes = [b.shape[0] for b, _ in train_ds]
  assert sizes == [3, 3, 2]                                # last batch of the epoch is smaller (no drop_remainder)
  assert sorted(sizes) == sorted(b.shape[0] for b, _ in train_ds)   # re-iterable across epochs
  val = np.concatenate([b for b, _ in val_ds])
  norm = (signals - signals.min()) / (signals.max() - signals.min())
  np.testing.assert_allclose(val, norm[8:], rtol=1e-6)
  assert val.min() >= 0.0 and val.max() <= 1.0


def _tf_example_class():
  """tf.train.Example built from tensorflow/core/example/{example,feature}.proto's published field numbers with the
  protobuf runtime (an implementation independent of this repo's reader / writer)."""
  from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
  F = descriptor_pb2.FieldDescriptorProto
  fdp = descriptor_pb2.FileDescriptorProto(name='tf_example_for_test.proto', package='tensorflow', syntax='proto3')

  def msg(name, fields=()):
    m = fdp.message_type.add(name=name)
    for fname, number, label, ftype, type_name in fields:
      m.field.add(name=fname, number=number, label=label, type=ftype, type_name=type_name)
    return m

  msg('BytesList', [('value', 1, F.LABEL_REPEATED, F.TYPE_BYTES, None)])
  msg('FloatList', [('value', 1, F.LABEL_REPEATED, F.TYPE_FLOAT, None)])
  msg('Int64List', [('value', 1, F.LABEL_REPEATED, F.TYPE_INT64, None)])
  feature = msg('Feature')
  feature.oneof_decl.add(name='kind')
  for number, (fname, tname) in enumerate((('bytes_list', 'BytesList'), ('float_list', 'FloatList'), ('int64_list', 'Int64List')), 1):
    feature.field.add(name=fname, number=number, label=F.LABEL_OPTIONAL, type=F.TYPE_MESSAGE, type_name='.tensorflow.' + tname,
                      oneof_index=0)
  features = msg('Features')
  entry = features.nested_type.add(name='FeatureEntry')
  entry.options.map_entry = True
  entry.field.add(name='key', number=1, label=F.LABEL_OPTIONAL, type=F.TYPE_STRING)
  entry.field.add(name='value', number=2, label=F.LABEL_OPTIONAL, type=F.TYPE_MESSAGE, type_name='.tensorflow.Feature')
  features.field.add(name='feature', number=1, label=F.LABEL_REPEATED, type=F.TYPE_MESSAGE,
                     type_name='.tensorflow.Features.FeatureEntry')
  msg('Example', [('features', 1, F.LABEL_OPTIONAL, F.TYPE_MESSAGE, '.tensorflow.Features')])
  pool = descriptor_pool.DescriptorPool()
  pool.Add(fdp)
  return message_factory.GetMessageClass(pool.FindMessageTypeByName('tensorflow.Example'))


def test_reader_parses_bytes_written_by_an_independent_protobuf_serializer(tmp_path):
  """dataset/generate_tfrecords.py:128-139 builds tf.train.Example(features{signal, spike: bytes_list}) and writes it
  with tf.io.TFRecordWriter. The payload here comes from the protobuf runtime (not from this repo's writer), carries
  an extra float_list / int64_list feature the reader must skip, and the frame is assembled by hand."""
  Example = _tf_example_class()
  rng = np.random.RandomState(3)
  sig, spk = rng.rand(16, 5).astype(np.float32), (rng.rand(16, 5) > 0.8).astype(np.float32)
  ex = Example()
  ex.features.feature['spike'].bytes_list.value.append(spk.tobytes())
  ex.features.feature['rate'].float_list.value.extend([1.5, 2.5])
  ex.features.feature['signal'].bytes_list.value.append(sig.tobytes())
  ex.features.feature['id'].int64_list.value.append(300)
  payload = ex.SerializeToString()
  got = D.parse_example(payload)
  np.testing.assert_array_equal(np.frombuffer(got['signal'], np.float32).reshape(16, 5), sig)
  np.testing.assert_array_equal(np.frombuffer(got['spike'], np.float32).reshape(16, 5), spk)
  assert 'rate' not in got and 'id' not in got
  # and the other direction: the repo's writer produces bytes the protobuf runtime parses to the same message
  back = Example.FromString(D.serialize_example(sig, spk))
  assert set(back.features.feature) == {'signal', 'spike'}
  assert back.features.feature['signal'].bytes_list.value[0] == sig.tobytes()
  # TFRecord frame: u64 length | masked crc32c(length) | payload | masked crc32c(payload)
  header = struct.pack('<Q', len(payload))
  path = str(tmp_path / 'train-001-of-001.record')
  with open(path, 'wb') as f:
    f.write(header + struct.pack('<I', D.masked_crc32c(header)) + payload + struct.pack('<I', D.masked_crc32c(payload)))
  assert list(D.read_records(path, verify_crc=True)) == [payload]


def test_hand_assembled_example_known_answer():
  """Byte-for-byte known answer from the protobuf wire format: Example{1: Features{1: entry{1: "signal", 2: Feature{1:
  BytesList{1: 01 02 03 04}}}}} = 0A 14 | 0A 12 | 0A 06 'signal' | 12 08 | 0A 06 | 0A 04 01 02 03 04."""
  raw = bytes([0x0A, 0x14, 0x0A, 0x12, 0x0A, 0x06]) + b'signal' + bytes([0x12, 0x08, 0x0A, 0x06, 0x0A, 0x04, 1, 2, 3, 4])
  got = D.parse_example(raw)
  assert set(got) == {'signal'} and bytes(got['signal']) == bytes([1, 2, 3, 4])
  # masked crc of a known frame header (length 20): standard crc32c, rotated right by 15 and offset 0xa282ead8
  c = D.crc32c(struct.pack('<Q', 20))
  assert D.masked_crc32c(struct.pack('<Q', 20)) == ((((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF)


def test_shuffle_buffer_order_is_a_windowed_permutation():
  """tf.data's shuffle(buffer_size) (dataset_helper.py:172): element i cannot be emitted before position i - buffer + 1"""
  rng = np.random.RandomState(0)
  for n, buf in ((50, 8), (50, 1), (50, 50), (50, 500), (7, 3), (1, 4)):
    order = D.shuffle_buffer_order(n, buf, rng)
    assert sorted(order.tolist()) == list(range(n))
    pos = np.empty(n, np.int64)
    pos[order] = np.arange(n)
    assert (pos >= np.arange(n) - buf + 1).all()
    if buf == 1:
      assert order.tolist() == list(range(n))
  # a full-size buffer is a uniform shuffle: every element reaches the first position with probability 1 / n
  first = np.bincount([D.shuffle_buffer_order(5, 5, rng)[0] for _ in range(4000)], minlength=5) / 4000.0
  assert np.abs(first - 0.2).max() < 0.03
  # a small buffer keeps the epoch roughly in file order
  assert np.abs(D.shuffle_buffer_order(1000, 10, rng) - np.arange(1000)).max() < 200


def test_rank_shards_are_disjoint_and_equal(tmp_path):
  for n, world in ((10, 2), (11, 4), (8, 8), (1100, 8)):
    shards = [D.shard_for_rank(n, r, world) for r in range(world)]
    assert len({len(s) for s in shards}) == 1 and len(shards[0]) == n // world
    allidx = np.concatenate(shards)
    assert len(set(allidx.tolist())) == len(allidx) and allidx.max() < n
  with pytest.raises(ValueError):
    D.shard_for_rank(3, 0, 4)
  # through get_dataset: two ranks, 9 training signals -> 4 each, the same number of steps, nothing shared
  rng = np.random.RandomState(2)
  signals = rng.rand(12, 32, 4).astype(np.float32)
  D.write_dataset(str(tmp_path), signals, np.zeros_like(signals), train_size=9, num_per_shard=4)
  seen = []
  for rank in range(2):
    hp = argparse.Namespace(input_dir=str(tmp_path), output_dir=str(tmp_path / 'runs'), batch_size=3, noise_dim=4,
                            rank=rank, world_size=2)
    train_ds, val_ds = D.get_dataset(hp)
    assert hp.train_size == 4 and hp.train_steps == 2 and hp.validation_size == 3
    assert [b.shape[0] for b, _ in train_ds] == [3, 1]
    seen.append(np.concatenate([b for b, _ in train_ds]))
    assert sum(b.shape[0] for b, _ in val_ds) == 3          # validation is not sharded
  rows = {tuple(np.round(x.reshape(-1)[:6], 6)) for x in np.concatenate(seen)}
  assert len(rows) == 8


def test_validation_cache_written_once_with_raw_scale_signals_and_int8_spikes(tmp_path):
  """dataset_helper.py:12-31,196-197: only with --save_generated"""
  from calciumgan_b200.utils import h5_helper
  rng = np.random.RandomState(3)
  signals = rng.rand(9, 32, 4).astype(np.float32) * 5 - 2
  spikes = (rng.rand(9, 32, 4) > 0.8).astype(np.float32)
  data_dir = str(tmp_path / 'tfrecords')
  D.write_dataset(data_dir, signals, spikes, train_size=5, num_per_shard=4)
  hp = argparse.Namespace(input_dir=data_dir, output_dir=str(tmp_path / 'runs'), batch_size=3, noise_dim=4, save_generated='')
  D.get_dataset(hp)
  assert not h5_helper.exists(hp.validation_cache)
  for _ in range(2):        # the second call finds the cache and leaves it alone
    hp = argparse.Namespace(input_dir=data_dir, output_dir=str(tmp_path / 'runs'), batch_size=3, noise_dim=4, save_generated='last')
    train_ds, val_ds = D.get_dataset(hp)
    assert hp.validation_cache == os.path.join(hp.output_dir, 'generated', 'validation.h5')
    got = h5_helper.get(hp.validation_cache, 'signals')
    assert got.shape == (4, 32, 4) and got.dtype == np.float32
    np.testing.assert_allclose(got, signals[5:], rtol=1e-5, atol=1e-5)
    sp = h5_helper.get(hp.validation_cache, 'spikes')
    assert sp.dtype == np.int8
    np.testing.assert_array_equal(sp, spikes[5:].astype(np.int8))
  # batches carry the spikes like the reference's (signal, spike) pairs
  s0, k0 = next(iter(val_ds))
  assert k0.shape == s0.shape and set(np.unique(k0)) <= {0.0, 1.0}


def test_surrogate_dataset(tmp_path):
  """dataset_helper.py:53-110: training.pkl holds (trials, neurons, time); 8192 train, the rest validate"""
  rng = np.random.RandomState(4)
  raw = (rng.rand(8200, 3, 16) * 4 - 1).astype(np.float32)
  spikes = (rng.rand(8200, 16, 3) > 0.9).astype(np.int8)
  with open(str(tmp_path / 'training.pkl'), 'wb') as file:
    pickle.dump({'signals': raw, 'spikes': spikes}, file)
  hp = argparse.Namespace(input_dir=str(tmp_path), output_dir=str(tmp_path / 'runs'), batch_size=1024, noise_dim=4, surrogate_ds=True)
  train_ds, val_ds = D.get_dataset(hp)
  assert hp.signal_shape == (16, 3) and hp.spike_shape == (16, 3) and hp.num_neurons == 3 and hp.sequence_length == 16
  assert hp.train_size == 8192 and hp.validation_size == 8 and hp.train_steps == 8 and hp.validation_steps == 1
  assert hp.normalize and abs(hp.signals_min - raw.min()) < 1e-6 and abs(hp.signals_max - raw.max()) < 1e-6
  val = np.concatenate([b for b, _ in val_ds])
  np.testing.assert_allclose(val, (raw[8192:].transpose(0, 2, 1) - raw.min()) / (raw.max() - raw.min()), rtol=1e-5, atol=1e-6)
  assert train_ds.buffer_size == 2048 and sum(b.shape[0] for b, _ in train_ds) == 8192
