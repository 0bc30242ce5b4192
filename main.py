"""Training driver with the reference's command line (main.py:227-267 of bryanlimy/CalciumGAN).

Same flags and defaults; the epoch loop calls `gan.train(signal)` exactly like main.py:33-75, logs the reference's
TensorBoard scalar tags and checkpoints with the reference's pickle layout. Signals come from, in this order:
the reference's TFRecord layout (`<input_dir>/info.pkl` + `train-*.record` / `validation-*.record`, read without
TensorFlow), `<input_dir>/signals.npy` (float32 (N, seq, neurons), already in [0, 1]), or with --synthetic a seeded
uniform generator of shape (N, 2048, 102). Spike analysis / plotting of the reference are out of scope.
"""
import argparse
import os
import sys
from shutil import rmtree
from time import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def init_distributed(hparams):
  """Batch-sharded data parallelism (no reference counterpart): under `torchrun --nproc-per-node N main.py ...`
  every rank binds its GPU and joins the NCCL group before the engine is built, trains on its own shard of the training
  split (`dataset_helper.shard_for_rank`) and exchanges gradients inside `gan.train`; rank 0 alone writes summaries,
  checkpoints and generated signals. A plain `python main.py` leaves world_size = 1."""
  hparams.world_size, hparams.rank = 1, 0
  if int(os.environ.get('WORLD_SIZE', '1')) > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if not dist.is_initialized():
      dist.init_process_group('nccl')
    hparams.world_size, hparams.rank = dist.get_world_size(), dist.get_rank()


class NullSummary(object):
  """what ranks other than 0 log to: every Summary method is accepted and dropped"""

  def __getattr__(self, name):
    return lambda *args, **kwargs: None


def get_dataset(hparams):
  """Sets the hparams fields the models read (dataset_helper.py:84-91,120-136,186)."""
  from calciumgan_b200.utils import dataset_helper
  if getattr(hparams, 'surrogate_ds', False) or (not hparams.synthetic and os.path.exists(os.path.join(hparams.input_dir, 'info.pkl'))):
    train_ds, validation_ds = dataset_helper.get_dataset(hparams)
    return (lambda: iter(train_ds)), (lambda: iter(validation_ds))
  path = os.path.join(hparams.input_dir, 'signals.npy')
  if hparams.synthetic or not os.path.exists(path):
    rng = np.random.RandomState(1234)
    signals = rng.uniform(0, 1, size=(hparams.synthetic_size, 2048, 102)).astype(np.float32)
  else:
    signals = np.load(path).astype(np.float32)
  n_train = int(len(signals) * 0.9) if len(signals) >= 10 else len(signals)
  train, val = signals[:n_train], signals[n_train:] if n_train < len(signals) else signals[:1]
  if getattr(hparams, 'world_size', 1) > 1:
    train = train[dataset_helper.shard_for_rank(len(train), hparams.rank, hparams.world_size)]
  hparams.train_size, hparams.validation_size = len(train), len(val)
  hparams.signal_shape = tuple(train.shape[1:])
  hparams.sequence_length, hparams.num_neurons = train.shape[1], train.shape[-1]
  hparams.num_channels = train.shape[-1]
  hparams.normalize, hparams.fft, hparams.conv2d = True, False, False
  hparams.signals_min, hparams.signals_max = 0.0, 1.0
  hparams.noise_shape = (hparams.noise_dim,)
  hparams.train_steps = int(np.ceil(len(train) / hparams.batch_size))
  hparams.validation_steps = int(np.ceil(len(val) / hparams.batch_size))
  dataset_helper.set_generated_dir(hparams)

  def batches(x, shuffle):
    idx = np.random.permutation(len(x)) if shuffle else np.arange(len(x))
    for i in range(0, len(x), hparams.batch_size):    # no drop_remainder (dataset_helper.py:173)
      yield x[idx[i:i + hparams.batch_size]], None

  return (lambda: batches(train, True)), (lambda: batches(val, False))


def _train_batches(hparams, train_ds, cache):
  """Device batches of one epoch. With the device-resident cache (the reference's `train_ds.cache()`,
  dataset_helper.py:171, kept in HBM) the first epoch uploads and caches every batch, later epochs only send shuffled
  indices; otherwise every batch is streamed from pinned host memory with the copy of batch i+1 under step i."""
  from calciumgan_b200.utils.prefetch import prefetch_to_device
  if cache is None:
    return prefetch_to_device(train_ds())
  if not cache.complete:
    cache.filled = 0
    return cache.fill_from(train_ds())
  return cache.batches(hparams.batch_size, shuffle=True)


def train(hparams, train_ds, gan, summary, epoch, cache=None):
  gen_losses, dis_losses, gradient_penalties = [], [], []
  start = time()
  batch_count = 0
  for signal, _ in _train_batches(hparams, train_ds, cache):
    if hparams.profile and batch_count == 2 and epoch == 1:
      summary.profiler_trace()      # main.py:45-47: the 2nd batch of the 2nd epoch opens the profiled window
    gen_loss, dis_loss, gradient_penalty, metrics = gan.train(signal)
    if hparams.profile and batch_count == 6 and epoch == 1:
      summary.profiler_export()     # main.py:51-52
    batch_count += 1
    gen_losses.append(gen_loss)
    dis_losses.append(dis_loss)
    if gradient_penalty is not None:
      gradient_penalties.append(gradient_penalty)
    hparams.global_step += 1
  end = time()
  gen_loss, dis_loss = np.mean(gen_losses), np.mean(dis_losses)
  summary.log(gen_loss, dis_loss, np.mean(gradient_penalties) if gradient_penalties else None, elapse=end - start,
              gan=gan, step=epoch, training=True)
  if hparams.verbose:
    print('Train epoch {:03d}: generator {:.4f} discriminator {:.4f} gradient_penalty {:.4f} elapse {:.2f}s'.format(
        epoch, gen_loss, dis_loss, np.mean(gradient_penalties), end - start))
  return gen_loss, dis_loss


def validate(hparams, validation_ds, gan, summary, epoch):
  from calciumgan_b200.utils import utils
  start = time()
  gen_losses, dis_losses, gradient_penalties, results = [], [], [], {}
  save_generated = utils.save_generated_at(hparams, epoch) and getattr(hparams, 'rank', 0) == 0      # main.py:81-84
  for signal, _ in validation_ds():
    fake, gen_loss, dis_loss, gradient_penalty, metrics = gan.validate(signal)
    if save_generated:
      utils.save_fake_signals(hparams, epoch, signals=fake)     # main.py:105-106
    gen_losses.append(gen_loss)
    dis_losses.append(dis_loss)
    gradient_penalties.append(gradient_penalty)
    for k, v in metrics.items():
      results.setdefault(k, []).append(v)
  results = {key: np.mean(item) for key, item in results.items()}
  summary.log(np.mean(gen_losses), np.mean(dis_losses), np.mean(gradient_penalties), metrics=results,
              elapse=time() - start, gan=gan, step=epoch, training=False)
  if hparams.verbose:
    print('Validation epoch {:03d}: generator {:.4f} discriminator {:.4f} gradient_penalty {:.4f} '.format(
        epoch, np.mean(gen_losses), np.mean(dis_losses), np.mean(gradient_penalties)) +
          ' '.join('{} {:.5f}'.format(k.split('/')[-1], v) for k, v in results.items()))
  return results


def main(hparams, return_metrics=False):
  from calciumgan_b200.algorithms.registry import get_algorithm
  from calciumgan_b200.models.registry import get_models
  from calciumgan_b200.utils import utils

  init_distributed(hparams)
  chief = hparams.rank == 0
  if chief and hparams.clear_output_dir and os.path.exists(hparams.output_dir):
    rmtree(hparams.output_dir)
  os.makedirs(hparams.output_dir, exist_ok=True)
  if hparams.world_size > 1:
    import torch.distributed as dist
    dist.barrier()                 # nobody reads output_dir (checkpoints to resume from) before rank 0 has cleared it
  np.random.seed(1234 + hparams.rank)

  from calciumgan_b200.utils.summary_helper import Summary
  summary = Summary(hparams) if chief else NullSummary()
  train_ds, validation_ds = get_dataset(hparams)
  generator, discriminator = get_models(hparams, summary)
  gan = get_algorithm(hparams, generator, discriminator, summary)
  utils.load_models(hparams, gan)
  if chief:
    utils.save_hparams(hparams)     # main.py:184 of the reference

  cache = None
  if not hparams.no_device_cache:
    from calciumgan_b200.utils.dataset_cache import DeviceDatasetCache
    if DeviceDatasetCache.fits(hparams.train_size, hparams.signal_shape):
      cache = DeviceDatasetCache(gan.engine, hparams.train_size, hparams.signal_shape)

  start = time()
  results = {}
  for epoch in range(hparams.start_epoch, hparams.epochs):
    train(hparams, train_ds, gan, summary, epoch, cache)
    results = validate(hparams, validation_ds, gan, summary, epoch)
    if chief and not hparams.skip_checkpoints and (epoch % 10 == 0 or epoch == hparams.epochs - 1):
      utils.save_models(hparams, gan, epoch)
  summary.scalar('elapse/total', time() - start)
  summary.flush()
  if chief and getattr(hparams, 'surrogate_ds', False):    # main.py:219-221: samples for the surrogate metrics
    utils.generate_dataset(hparams, gan=gan, num_samples=2 * 10**6)
  if hparams.verbose:
    print('elapse/total {:.2f}s'.format(time() - start))
  if return_metrics:
    return results


# The reference's command line (main.py:229-261): (flag, default, help). A bool default makes a store_true switch, any other
# default fixes the value type; `choices` are listed separately.
REFERENCE_FLAGS = (
    ('input_dir', 'dataset/tfrecords', 'TFRecord shards + info.pkl'),
    ('output_dir', 'runs', 'TensorBoard events, hparams.json, checkpoints/, generated/'),
    ('batch_size', 64, None),
    ('num_units', 32, 'channel multiplier of both networks'),
    ('kernel_size', 24, None),
    ('strides', 2, None),
    ('m', 2, 'phase shuffle m'),
    ('n', 2, 'phase shuffle n'),
    ('epochs', 20, None),
    ('dropout', 0.2, 'unused by the calciumgan model'),
    ('learning_rate', 0.0001, None),
    ('noise_dim', 32, None),
    ('gradient_penalty', 10.0, 'WGAN-GP lambda'),
    ('model', 'wavegan', None),
    ('activation', 'leakyrelu', None),
    ('batch_norm', False, None),
    ('layer_norm', False, None),
    ('algorithm', 'wgan-gp', None),
    ('n_critic', 5, 'number of steps between each generator update'),
    ('clear_output_dir', False, None),
    ('save_generated', '', 'save the generated validation signals: after the last epoch, or after every 10th as well'),
    ('plot_weights', False, 'accepted for compatibility; plotting is out of scope'),
    ('skip_checkpoints', False, None),
    ('mixed_precision', False, 'bf16 tensor-core path'),
    ('profile', False, 'open a cudaProfilerStart/Stop + NVTX window around batches 2-6 of the 2nd epoch'),
    ('dpi', 120, 'accepted for compatibility; plotting is out of scope'),
    ('verbose', 1, None),
)
# additions (not in the reference): data source when no TFRecord pipeline is available, and the dataset cache switch
EXTRA_FLAGS = (
    ('synthetic', False, 'uniform [0,1) signals of shape (N, 2048, 102)'),
    ('synthetic_size', 512, None),
    ('no_device_cache', False, 'stream every batch from host memory instead of caching the training set in HBM'),
)
FLAG_CHOICES = {'save_generated': ['', 'last', 'all']}


def build_parser():
  parser = argparse.ArgumentParser()
  for name, default, text in REFERENCE_FLAGS + EXTRA_FLAGS:
    if isinstance(default, bool):
      parser.add_argument('--' + name, action='store_true', help=text)
    else:
      parser.add_argument('--' + name, default=default, type=type(default), choices=FLAG_CHOICES.get(name), help=text)
  return parser


if __name__ == '__main__':
  params = build_parser().parse_args()
  params.global_step = 0
  params.surrogate_ds = 'surrogate' in params.input_dir      # main.py:265 of the reference
  main(params)
