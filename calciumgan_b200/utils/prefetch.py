"""Host -> device input prefetch: the next batch's H2D copy runs on a side stream while the current step computes.

Stands in for the reference's `tf.data` `prefetch(4)` (gan/utils/dataset_helper.py:174,181): `gan.train(signal)`
receives a device tensor whose copy is ordered before its first use (the consumer stream waits on the copy event).
"""
import numpy as np
import torch


def prefetch_to_device(batches, device=None, depth=2):
  """Yield (cuda_signal, extra) for every item of `batches` ((signal, extra) pairs or bare signals; numpy or torch,
  float32). `depth` device buffers (+ pinned staging buffers when the source is not pinned) are reused; batches may
  be smaller than the first one (ragged last batch of an epoch)."""
  device = torch.device('cuda', torch.cuda.current_device()) if device is None else device
  copy_stream = torch.cuda.Stream(device=device)
  it = iter(batches)
  bufs = [dict(dev=None, pinned=None, free=None, ready=None) for _ in range(depth)]
  queue = []
  state = dict(slot=0)

  def issue():
    try:
      item = next(it)
    except StopIteration:
      return
    signal, extra = item if isinstance(item, tuple) else (item, None)
    src = signal if isinstance(signal, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(signal, np.float32))
    src = src.contiguous()
    buf = bufs[state['slot']]
    slot = state['slot']
    state['slot'] = (slot + 1) % depth
    n = src.shape[0]
    if buf['dev'] is None or buf['dev'].shape[1:] != src.shape[1:] or buf['dev'].shape[0] < n:
      buf['dev'] = torch.empty(src.shape, dtype=torch.float32, device=device)
      buf['pinned'] = None
      # the block may be recycled from main-stream work still in flight, and is written on copy_stream from now on
      copy_stream.wait_stream(torch.cuda.current_stream())
      buf['dev'].record_stream(copy_stream)
    host = src
    if not src.is_pinned():
      if buf['pinned'] is None or buf['pinned'].shape[0] < n:
        buf['pinned'] = torch.empty(buf['dev'].shape, dtype=torch.float32).pin_memory()
      if buf['ready'] is not None:
        buf['ready'].synchronize()             # the previous async copy OUT of this staging buffer has finished
      buf['pinned'][:n].copy_(src)
      host = buf['pinned'][:n]
    if buf['free'] is not None:
      copy_stream.wait_event(buf['free'])      # the step that consumed this buffer has finished
    with torch.cuda.stream(copy_stream):
      buf['dev'][:n].copy_(host, non_blocking=True)
      ready = torch.cuda.Event()
      ready.record(copy_stream)
    buf['ready'] = ready
    queue.append((slot, n, ready, extra))

  for _ in range(depth):
    issue()
  while queue:
    slot, n, ready, extra = queue.pop(0)
    torch.cuda.current_stream().wait_event(ready)
    yield bufs[slot]['dev'][:n], extra
    free = torch.cuda.Event()
    free.record(torch.cuda.current_stream())
    bufs[slot]['free'] = free
    issue()
