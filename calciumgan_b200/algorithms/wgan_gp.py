"""WGAN-GP algorithm plugin: the reference's train-step interface over the CUDA engine.

Mirrors gan/algorithms/wgan_gp.py:9-95: `train(inputs)` runs n_critic critic updates on the
same batch plus one generator update and returns (gen_loss, dis_loss, gradient_penalty,
metrics). Optional keyword arguments inject the random draws (noise, interpolation alpha,
phase-shuffle shifts) so the step can be checked against the reference semantics.

Data parallelism (no reference counterpart): with torch.distributed initialised,
each rank runs its batch shard; the flat fp32 gradient buffer is all-reduced over NCCL before
the fused Adam (which folds in 1/world_size). PhaseShuffle shifts are per-call scalars shared
by the whole batch, so all ranks draw them from the same seeded stream.
"""
import os

import numpy as np
import torch

from .registry import register
from .gan import GAN, metrics_from_scalars
from .. import _lib as L


def _dist():
  import torch.distributed as dist
  return dist if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else None


@register('wgan-gp')
class WGAN_GP(GAN):

  def __init__(self, hparams, generator, discriminator, summary=None):
    super().__init__(hparams, generator, discriminator, summary)

    self.penalty = hparams.gradient_penalty
    self.n_critic = hparams.n_critic
    self.conv2d = getattr(hparams, 'conv2d', False)
    if self.n_critic != self.engine.cfg.n_critic:
      raise ValueError('n_critic changed after the models were built')

  # ------------------------------------------------------------------ losses (wgan_gp.py:19-62)
  def generator_loss(self, fake_output):
    return -torch.mean(fake_output)

  def interpolation(self, real, fake, alpha=None):
    real, fake = self.engine.to_device(real), self.engine.to_device(fake)
    if alpha is None:
      alpha = torch.rand((real.shape[0], 1, 1), device=real.device)
    alpha = self.engine.to_device(alpha).reshape(-1, 1, 1)
    return (alpha * real) + ((1 - alpha) * fake)

  def gradient_penalty(self, real, fake, training=True, alpha=None, shifts=None):
    interpolated = self.interpolation(real, fake, alpha)
    if shifts is None:
      m = self.engine.cfg.phase_m
      shifts = np.random.randint(-m, m + 1, size=4)
    _, sumsq = self.engine.gp_debug(interpolated, shifts)
    return torch.mean(torch.square(torch.sqrt(sumsq) - 1.0))

  def discriminator_loss(self, real_output, fake_output, real=None, fake=None, training=True):
    real_loss = -torch.mean(real_output)
    fake_loss = torch.mean(fake_output)
    gradient_penalty = self.gradient_penalty(real, fake, training=training)
    loss = real_loss + fake_loss + self.penalty * gradient_penalty
    return loss, gradient_penalty

  # ------------------------------------------------------------------ steps (wgan_gp.py:22-36,64-95)
  # ---------------------------------------------------------------- data-parallel gradient exchange
  def _peer_setup(self, dist):
    """Gradient buffers in symmetric memory (every rank maps every peer's buffer over NVLink): the exchange is then ONE
    small kernel of this library (cg_reduce_peer_grads) between two cross-rank barriers instead of an NCCL all-reduce.
    Returns False (NCCL path) when symmetric memory is unavailable or CG_DP_COMM=nccl."""
    if hasattr(self, '_peer'):
      return self._peer is not None
    self._peer = None
    eng = self.engine
    world = dist.get_world_size()
    # measured same-box: the peer-memory exchange wins at 2 ranks (12.25 vs 12.46 ms per step), NCCL at 4
    # (12.63-12.69 vs 12.78); CG_DP_COMM=p2p|nccl overrides
    mode = os.environ.get('CG_DP_COMM') or ('p2p' if world == 2 else 'nccl')
    if eng.device.type != 'cuda' or mode != 'p2p' or world not in (2, 4, 8):
      return False
    try:
      import torch.distributed._symmetric_memory as symm
      group = dist.group.WORLD
      peer = {}
      for which in (L.GENERATOR, L.DISCRIMINATOR):
        n = eng.num_params(which)
        buf = symm.empty((n + 3) // 4 * 4, dtype=torch.float32, device=eng.device)
        buf.zero_()
        hdl = symm.rendezvous(buf, group)
        # the tensor may sit at an offset inside a pooled symmetric block: same offset on every rank
        delta = buf.data_ptr() - int(hdl.buffer_ptrs[dist.get_rank()])
        ptrs = [int(hdl.buffer_ptrs[r]) + delta for r in range(world)]
        red, rptrs = None, None
        if world > 2:    # two-phase exchange: the reduced gradient is peer-mapped too
          red = symm.empty((n + 7) // 4 * 4, dtype=torch.float32, device=eng.device)
          red.zero_()
          rh = symm.rendezvous(red, group)
          rdelta = red.data_ptr() - int(rh.buffer_ptrs[dist.get_rank()])
          rptrs = [int(rh.buffer_ptrs[r]) + rdelta for r in range(world)]
        peer[which] = (buf, hdl, ptrs, red, rptrs)
      torch.cuda.synchronize()
      dist.barrier()
      for which, (buf, _, _, red, _) in peer.items():
        eng.set_grad_buffer(which, buf)
        if red is not None:
          eng.set_reduced_buffer(which, red)
      self._peer = peer
    except Exception as e:   # noqa: BLE001 -- any failure here just means "use NCCL"
      if dist.get_rank() == 0:
        print('calciumgan_b200: peer-memory gradient exchange unavailable (%s: %s); using NCCL all-reduce' % (type(e).__name__, e))
    # every rank must take the same path
    flag = torch.tensor([1.0 if self._peer is not None else 0.0], device=eng.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if float(flag.item()) == 0.0 and self._peer is not None:
      for which in self._peer:
        eng.set_grad_buffer(which, None)
        eng.set_reduced_buffer(which, None)
      self._peer = None
    return self._peer is not None

  def _allreduce_start(self, dist, which):
    """Bucketed gradient all-reduce overlapped with the backward pass: bucket b is reduced on a side stream as soon as
    the library's event for its last writer has fired; the remaining wgrad kernels keep running on the main stream.
    Returns a handle for _allreduce_finish; nothing on the main stream waits yet, so the caller can enqueue work that
    does not need the reduced gradients (the next sub-step's generator forward) in between."""
    eng = self.engine
    if eng.device.type != 'cuda':        # host-logic tests drive this class over a CPU stub engine (gloo)
      for view, _ in eng.grad_buckets(which):
        dist.all_reduce(view)
      return None
    if not hasattr(self, '_comm_stream'):
      self._comm_stream = torch.cuda.Stream(device=eng.device)
    if os.environ.get('CG_DP_COMM') == 'none':     # timing experiment only: no exchange at all (replicas diverge)
      return None
    if self._peer_setup(dist):
      buf, hdl, ptrs, red, rptrs = self._peer[which]
      rank = dist.get_rank()
      with torch.cuda.stream(self._comm_stream):
        eng.stream_wait_bucket(which, eng.num_buckets(which) - 1, self._comm_stream)   # the last writer of this model's gradients
        hdl.barrier(channel=0)                      # every rank's gradients are complete (and its last gather is done)
        if red is None:                             # 2 ranks: one kernel pulls the peer's buffer
          eng.reduce_peer_grads(which, ptrs, self._comm_stream)
          hdl.barrier(channel=1)                    # every rank has finished reading: buffers may be overwritten
        else:                                       # 4 / 8 ranks: reduce-scatter, barrier, all-gather
          eng.peer_reduce_scatter(which, ptrs, rank, self._comm_stream)
          hdl.barrier(channel=1)                    # every slice is complete; gradient buffers may be overwritten
          eng.peer_all_gather(which, rptrs, rank, self._comm_stream)
      return 'peer'
    works = []
    with torch.cuda.stream(self._comm_stream):
      buckets = eng.grad_buckets(which)
      if os.environ.get('CG_DP_BUCKETS') == '1':   # experiment: one all-reduce per model after the last writer
        eng.stream_wait_bucket(which, buckets[-1][1], self._comm_stream)
        works.append(dist.all_reduce(eng.grad_tensor(which), async_op=True))
        return works
      for view, b in buckets:
        eng.stream_wait_bucket(which, b, self._comm_stream)
        works.append(dist.all_reduce(view, async_op=True))
    return works

  def _allreduce_finish(self, works, which):
    """Make the main stream wait for the exchange, then Adam on the exchanged gradients."""
    eng = self.engine
    if works is None:
      eng.apply_update(which)
      return
    if works == 'peer':
      torch.cuda.current_stream().wait_stream(self._comm_stream)
      eng.apply_update_reduced(which)
      return
    with torch.cuda.stream(self._comm_stream):
      for w in works:
        w.wait()
    torch.cuda.current_stream().wait_stream(self._comm_stream)
    eng.apply_update(which)

  def _allreduce_buckets(self, dist, which):
    self._allreduce_finish(self._allreduce_start(dist, which), which)

  def _train_discriminator(self, inputs, noise=None, alpha=None, shifts=None):
    dist = _dist()
    if dist is None:
      s = self.engine.critic_step(inputs, noise, alpha, shifts, update=True)
    else:
      self._peer_setup(dist)     # before the step: it decides which buffer the gradients are accumulated into
      s = self.engine.critic_step(inputs, noise, alpha, shifts, update=False)
      self._allreduce_buckets(dist, L.DISCRIMINATOR)
    return float(s[L.S_DIS_LOSS]), float(s[L.S_GP])

  def _train_generator(self, inputs, noise=None, shifts=None):
    dist = _dist()
    if dist is None:
      s = self.engine.generator_step(inputs, noise, shifts, update=True)
    else:
      self._peer_setup(dist)
      s = self.engine.generator_step(inputs, noise, shifts, update=False)
      self._allreduce_buckets(dist, L.GENERATOR)
    return float(s[L.S_GEN_LOSS]), metrics_from_scalars(s)

  def train(self, inputs, noise=None, alpha=None, shifts=None):
    """wgan_gp.py:82-95. noise (n_critic+1, B, nd), alpha (n_critic, B), shifts (12*n_critic+4,)."""
    dist = _dist()
    if dist is None:
      s = self.engine.train_step(inputs, noise, alpha, shifts)
      return (float(s[L.S_GEN_LOSS]), float(s[L.S_DIS_LOSS]), float(s[L.S_GP]), metrics_from_scalars(s))
    return self._train_dp(dist, inputs, noise, alpha, shifts)

  def _train_dp(self, dist, inputs, noise, alpha, shifts):
    """One data-parallel train step. Per sub-step: backward with the bucketed all-reduce running behind it, then --
    before anything waits for the reduced gradients -- the generator part of the NEXT sub-step (it reads no critic
    weight: wgan_gp.py:65-66 / :23-26), so the all-reduce tail and the critic's Adam hide under ~0.3 ms of generator
    GEMMs instead of idling the GPU."""
    eng, nc = self.engine, self.n_critic
    self._peer_setup(dist)       # before the first sub-step: it decides which buffer the gradients are accumulated into
    real = eng.to_device(inputs)
    shifts = None if shifts is None else np.asarray(shifts, np.int32).reshape(-1)
    hist = torch.zeros((nc + 1, L.NUM_SCALARS), device=eng.device)
    scal = eng.scalars_tensor()
    overlap = getattr(eng, 'prefetch_generator', None) is not None and not getattr(self, 'no_dp_overlap', False)
    prefetched = False
    for i in range(nc):
      eng.critic_step(real, None if noise is None else noise[i], None if alpha is None else alpha[i],
                      None if shifts is None else shifts[12 * i:12 * i + 12], update=False, sync=False, same_real=i > 0,
                      want_fake32=False,   # the generator step below rewrites the fp32 output before anything reads it
                      gen_prefetched=prefetched)
      hist[i].copy_(scal)
      works = self._allreduce_start(dist, L.DISCRIMINATOR)
      if overlap:
        if i + 1 < nc:
          eng.prefetch_generator(real, None if noise is None else noise[i + 1], None if alpha is None else alpha[i + 1],
                                 for_generator_step=False, want_fake32=False)
        else:
          eng.prefetch_generator(real, None if noise is None else noise[nc], None, for_generator_step=True)
        prefetched = True
      self._allreduce_finish(works, L.DISCRIMINATOR)
    eng.generator_step(real, None if noise is None else noise[nc],
                       None if shifts is None else shifts[12 * nc:12 * nc + 4], update=False, sync=False,
                       gen_prefetched=prefetched)
    hist[nc].copy_(scal)
    self._allreduce_buckets(dist, L.GENERATOR)
    out = torch.zeros(L.NUM_SCALARS, device=eng.device)
    out[L.S_DIS_LOSS] = hist[:nc, L.S_DIS_LOSS].mean()
    out[L.S_GP] = hist[:nc, L.S_GP].mean()
    out[L.S_GEN_LOSS:] = hist[nc, L.S_GEN_LOSS:]
    dist.all_reduce(out)
    s = (out / dist.get_world_size()).cpu().numpy()
    return (float(s[L.S_GEN_LOSS]), float(s[L.S_DIS_LOSS]), float(s[L.S_GP]), metrics_from_scalars(s))

  def validate(self, inputs, noise=None, alpha=None, shifts=None):
    """gan.py:87-90 -> (fake, gen_loss, dis_loss, gradient_penalty, metrics)."""
    fake, s = self.engine.validate(inputs, noise, alpha, shifts)
    return (fake, float(s[L.S_GEN_LOSS]), float(s[L.S_DIS_LOSS]), float(s[L.S_GP]), metrics_from_scalars(s))
