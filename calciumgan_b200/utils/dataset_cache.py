"""Device-resident dataset cache: the reference's `train_ds.cache()` (gan/utils/dataset_helper.py:171) in HBM.

The reference reads its TFRecords once, caches the parsed signals and draws shuffled batches from the cache every epoch.
Here the cache lives on the GPU (a 180 GB part holds ~200k paper-size samples): the first pass uploads every batch once
(pinned staging, copies overlapped with the training step), later epochs send only a shuffled index vector (8 bytes per
sample) and assemble the batch with one gather kernel (cg_gather_rows). With 8 ranks sharing one host that removes
8 x 107 MB of pinned-host -> device traffic per step (the end-to-end limiter of round 1).
"""
import numpy as np
import torch


class DeviceDatasetCache(object):

  def __init__(self, engine, num_samples, sample_shape):
    self.engine = engine
    self.shape = tuple(int(s) for s in sample_shape)
    self.data = torch.empty((int(num_samples),) + self.shape, dtype=torch.float32, device=engine.device)
    self.filled = 0
    self.h2d_bytes = 0          # bytes copied host -> device so far (bench accounting)
    self._idx_pinned = None

  @staticmethod
  def fits(num_samples, sample_shape, device=None, fraction=0.5):
    """True when the cache takes at most `fraction` of the memory that is free right now."""
    free, _ = torch.cuda.mem_get_info(device)
    return int(num_samples) * int(np.prod(sample_shape)) * 4 <= fraction * free

  @property
  def complete(self):
    return self.filled == self.data.shape[0]

  def append(self, batch):
    """Store one uploaded batch (a CUDA tensor, e.g. from prefetch_to_device) and return the cached view of it."""
    n = batch.shape[0]
    dst = self.data[self.filled:self.filled + n]
    dst.copy_(batch)
    self.filled += n
    self.h2d_bytes += n * int(np.prod(self.shape)) * 4
    return dst

  def fill_from(self, host_batches):
    """First epoch: yield every batch while it is being cached (host batches: numpy / torch, (signal, extra) or bare)."""
    from .prefetch import prefetch_to_device
    for signal, extra in prefetch_to_device(host_batches, device=self.engine.device):
      yield self.append(signal), extra

  def batches(self, batch_size, shuffle=True, rng=None, drop_remainder=False):
    """Later epochs: shuffled batches assembled on the device; only the index vector crosses PCIe."""
    assert self.complete, 'cache is not complete yet'
    n = self.data.shape[0]
    rng = rng or np.random
    order = rng.permutation(n) if shuffle else np.arange(n)
    for i in range(0, n, batch_size):
      idx = order[i:i + batch_size]
      if drop_remainder and len(idx) < batch_size:
        return
      yield self.gather(idx), None

  def gather(self, idx):
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    if idx.min() < 0 or idx.max() >= self.filled:
      raise IndexError('sample index outside the cached range')
    if self._idx_pinned is None or self._idx_pinned[0].numel() < idx.size:
      cap = max(idx.size, 1024)
      self._idx_pinned = [torch.empty(cap, dtype=torch.int64).pin_memory() for _ in range(2)]
      self._idx_dev = [torch.empty(cap, dtype=torch.int64, device=self.engine.device) for _ in range(2)]
      self._idx_evt, self._idx_slot = [None, None], 0
    k = self._idx_slot
    self._idx_slot = 1 - k
    if self._idx_evt[k] is not None:
      self._idx_evt[k].synchronize()       # the copy that last read this staging buffer has finished
    self._idx_pinned[k][:idx.size].copy_(torch.from_numpy(idx))
    dev = self._idx_dev[k][:idx.size]
    dev.copy_(self._idx_pinned[k][:idx.size], non_blocking=True)
    self._idx_evt[k] = torch.cuda.Event()
    self._idx_evt[k].record()
    self.h2d_bytes += idx.size * 8
    return self.engine.gather_rows(self.data, dev)
