"""TF-free TFRecord reader/writer for the reference's dataset layout (dataset/generate_tfrecords.py,
gan/utils/dataset_helper.py): framing, protobuf payload, info.pkl -> hparams, batching without drop_remainder."""
import argparse
import struct

import numpy as np

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
  hp = argparse.Namespace(input_dir=str(tmp_path), batch_size=3, noise_dim=4)
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
