"""Scalar logging with the reference's TensorBoard tags (gan/utils/summary_helper.py:27-119,559-588).

Only the scalar part of the reference's `Summary` is reproduced (train / validation writers under output_dir,
`scalar`, `log`) plus the profiler hooks: `profiler_trace` / `profiler_export` (summary_helper.py:115-119, driven by
main.py:45-52 with --profile) open / close a CUDA profiler capture range (cudaProfilerStart / Stop, i.e. what
`ncu --profile-from-start off` and `nsys --capture-range=cudaProfilerApi` key on) and an NVTX range around the same
training batches the reference traces. Plotting and histograms are out of scope."""
import os


class Summary(object):

  def __init__(self, hparams, policy=None):
    from torch.utils.tensorboard import SummaryWriter
    self._hparams = hparams
    self.train_writer = SummaryWriter(hparams.output_dir)
    self.val_writer = SummaryWriter(os.path.join(hparams.output_dir, 'validation'))
    self.metrics_writer = SummaryWriter(os.path.join(hparams.output_dir, 'metrics'))
    self._policy = policy

  def _writer(self, training):
    return self.train_writer if training else self.val_writer

  def scalar(self, tag, value, step=0, training=True):
    self._writer(training).add_scalar(tag, float(value), global_step=step)

  def log(self, gen_loss, dis_loss, gradient_penalty, metrics=None, elapse=None, gan=None, step=0, training=True):
    """One epoch's scalars under the reference's tags (summary_helper.py:559-588): loss/generator, loss/discriminator,
    loss/gradient_penalty (WGAN-GP only), the signal metrics under their own tags, elapse, and -- validation with mixed
    precision -- model/loss_scale."""
    entries = [('loss/generator', gen_loss), ('loss/discriminator', dis_loss), ('loss/gradient_penalty', gradient_penalty)]
    entries += list((metrics or {}).items())
    entries.append(('elapse', elapse))
    if not training and gan is not None and getattr(self._hparams, 'mixed_precision', False):
      entries.append(('model/loss_scale', gan.gen_optimizer.loss_scale))
    for tag, value in entries:
      if value is not None:
        self.scalar(tag, value, step=step, training=training)

  def profiler_trace(self):
    """summary_helper.py:115-116 (tf.summary.trace_on): start of the profiled window."""
    import torch
    if torch.cuda.is_available():
      torch.cuda.synchronize()
      torch.cuda.profiler.start()
      torch.cuda.nvtx.range_push('calciumgan_b200/train')
    self._profiling = True

  def profiler_export(self):
    """summary_helper.py:118-119 (tf.summary.trace_export): end of the profiled window."""
    import torch
    if getattr(self, '_profiling', False) and torch.cuda.is_available():
      torch.cuda.synchronize()
      torch.cuda.nvtx.range_pop()
      torch.cuda.profiler.stop()
    self._profiling = False

  def flush(self):
    for w in (self.train_writer, self.val_writer, self.metrics_writer):
      w.flush()
