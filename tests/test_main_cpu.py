"""Host-side pieces of main.py that need no GPU: the command line (main.py:227-267 of the reference), the data sources,
rank sharding of the training split and what non-chief ranks log to."""
import os

import numpy as np

import main as driver


def _hp(tmp_path, *extra):
  hp = driver.build_parser().parse_args(['--input_dir', str(tmp_path / 'none'), '--output_dir', str(tmp_path / 'runs'),
                                         '--synthetic', '--synthetic_size', '40', '--batch_size', '8'] + list(extra))
  hp.global_step, hp.surrogate_ds = 0, False
  return hp


def test_reference_flags_and_defaults():
  hp = driver.build_parser().parse_args([])
  assert (hp.input_dir, hp.output_dir, hp.batch_size, hp.num_units, hp.kernel_size, hp.strides) == ('dataset/tfrecords', 'runs', 64, 32, 24, 2)
  assert (hp.m, hp.n, hp.epochs, hp.learning_rate, hp.noise_dim, hp.gradient_penalty) == (2, 2, 20, 0.0001, 32, 10.0)
  assert (hp.model, hp.activation, hp.algorithm, hp.n_critic, hp.save_generated) == ('wavegan', 'leakyrelu', 'wgan-gp', 5, '')
  assert not (hp.batch_norm or hp.layer_norm or hp.clear_output_dir or hp.skip_checkpoints or hp.mixed_precision or hp.profile)


def test_single_process_is_world_size_1(tmp_path, monkeypatch):
  monkeypatch.delenv('WORLD_SIZE', raising=False)
  hp = _hp(tmp_path)
  driver.init_distributed(hp)
  assert (hp.world_size, hp.rank) == (1, 0)
  train_ds, validation_ds = driver.get_dataset(hp)
  assert hp.train_size == 36 and hp.validation_size == 4 and hp.train_steps == 5 and hp.signal_shape == (2048, 102)
  assert hp.generated_dir == os.path.join(hp.output_dir, 'generated') and os.path.isdir(hp.generated_dir)
  assert [b.shape[0] for b, _ in train_ds()] == [8, 8, 8, 8, 4]            # no drop_remainder (dataset_helper.py:173)
  assert [b.shape[0] for b, _ in validation_ds()] == [4]


def test_synthetic_training_split_is_sharded_by_rank(tmp_path):
  shards = []
  for rank in range(4):
    hp = _hp(tmp_path)
    hp.world_size, hp.rank = 4, rank
    train_ds, validation_ds = driver.get_dataset(hp)
    assert hp.train_size == 9 and hp.train_steps == 2 and hp.validation_size == 4
    assert [b.shape[0] for b, _ in train_ds()] == [8, 1]                     # the same ragged last batch on every rank
    shards.append(np.concatenate([b for b, _ in train_ds()])[:, 0, 0])
  assert len(set(np.concatenate(shards).tolist())) == 36                     # disjoint shards


def test_null_summary_accepts_everything():
  s = driver.NullSummary()
  s.log(1.0, 2.0, None, metrics={'a': 1}, elapse=0.1, step=3, training=False)
  s.scalar('elapse/total', 1.0)
  s.flush()
  s.profiler_trace()


def test_summary_writes_the_reference_scalar_tags(tmp_path):
  """summary_helper.py:559-588: tags and writers (train events in output_dir, validation events in output_dir/validation)"""
  import argparse
  import glob
  from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
  from calciumgan_b200.utils.summary_helper import Summary
  hp = argparse.Namespace(output_dir=str(tmp_path), mixed_precision=True)
  s = Summary(hp)
  s.log(1.0, 2.0, None, metrics={'signals_metrics/min': 0.5}, elapse=3.0, step=4, training=True)
  gan = argparse.Namespace(gen_optimizer=argparse.Namespace(loss_scale=1.0))
  s.log(1.5, 2.5, 0.25, gan=gan, step=5, training=False)
  s.scalar('elapse/total', 9.0)
  s.flush()

  def scalars(directory):
    acc = EventAccumulator(directory)
    acc.Reload()
    return {tag: [(e.step, e.value) for e in acc.Scalars(tag)] for tag in acc.Tags()['scalars']}

  train = scalars(str(tmp_path))
  assert set(train) == {'loss/generator', 'loss/discriminator', 'signals_metrics/min', 'elapse', 'elapse/total'}
  assert train['loss/discriminator'] == [(4, 2.0)] and train['elapse'] == [(4, 3.0)]
  val = scalars(os.path.join(str(tmp_path), 'validation'))
  assert set(val) == {'loss/generator', 'loss/discriminator', 'loss/gradient_penalty', 'model/loss_scale'}
  assert val['loss/gradient_penalty'] == [(5, 0.25)]
  assert glob.glob(os.path.join(str(tmp_path), 'events.out.tfevents.*'))
