"""Host-side handle on one cg_ctx: device buffers, streams and tensor marshalling.

PyTorch is used here for device memory, streams and DLPack interchange only; every
arithmetic operation of the hot path runs inside libcalciumgan_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def hparams_to_config(hparams, max_batch=None, world_size=1, rank=0, force_simt=False, debug_flags=0):
  """Map the reference's hparams Namespace (main.py:229-261 + dataset_helper.py:84-91) to cg_config."""
  if getattr(hparams, 'batch_norm', False):
    raise NotImplementedError('--batch_norm couples samples and is out of scope (SURVEY §3.6-6)')
  if getattr(hparams, 'activation', 'leakyrelu') != 'leakyrelu':
    raise NotImplementedError('only --activation leakyrelu is implemented')
  if getattr(hparams, 'conv2d', False):
    raise NotImplementedError('conv2d signals are out of scope')
  cfg = L.CgConfig()
  cfg.seq_len = int(hparams.signal_shape[0])
  cfg.channels = int(getattr(hparams, 'num_channels', hparams.signal_shape[-1]))
  cfg.noise_dim = int(hparams.noise_dim)
  cfg.num_units = int(hparams.num_units)
  cfg.kernel_size = int(hparams.kernel_size)
  cfg.strides = int(hparams.strides)
  cfg.phase_m = int(hparams.m)
  cfg.layer_norm = int(bool(getattr(hparams, 'layer_norm', False)))
  cfg.normalize = int(bool(getattr(hparams, 'normalize', True)))
  cfg.max_batch = int(max_batch if max_batch is not None else hparams.batch_size)
  cfg.n_critic = int(getattr(hparams, 'n_critic', 5))
  cfg.precision = L.BF16 if getattr(hparams, 'mixed_precision', False) else L.FP32
  cfg.gp_lambda = float(getattr(hparams, 'gradient_penalty', 10.0))
  cfg.learning_rate = float(getattr(hparams, 'learning_rate', 1e-4))
  cfg.signals_min = float(getattr(hparams, 'signals_min', 0.0))
  cfg.signals_max = float(getattr(hparams, 'signals_max', 1.0))
  cfg.world_size = int(world_size)
  cfg.rank = int(rank)
  cfg.force_simt = int(bool(force_simt))
  cfg.debug_flags = int(debug_flags)
  return cfg


class _DevArray(object):
  """Expose a raw device pointer through __cuda_array_interface__ so torch can wrap it."""

  def __init__(self, ptr, n, typestr='<f4'):
    self.__cuda_array_interface__ = {
        'shape': (int(n),), 'typestr': typestr, 'data': (int(ptr), False), 'version': 2}


def phase_shuffle_index(w, shift):
  """Host copy of the kernels' PhaseShuffle index map (int32)."""
  idx = np.empty(w, np.int32)
  L.check(L.load().cg_phase_shuffle_index(int(w), int(shift), idx.ctypes.data_as(C.POINTER(C.c_int32))))
  return idx


class Engine(object):
  """One cg_ctx on the current CUDA device."""

  def __init__(self, cfg):
    self.lib = L.load()
    if not torch.cuda.is_available():
      raise RuntimeError('calciumgan_b200 needs a CUDA device (no CPU fallback)')
    self.device = torch.device('cuda', torch.cuda.current_device())
    torch.cuda.init()
    torch.zeros(1, device=self.device)   # make sure the primary context exists
    self.cfg = cfg
    ctx = C.c_void_p()
    L.check(self.lib.cg_create(C.byref(cfg), C.byref(ctx)))
    self.ctx = ctx
    self._scal = (C.c_float * L.NUM_SCALARS)()
    self._grad_views = {}

  def close(self):
    if getattr(self, 'ctx', None):
      self.lib.cg_destroy(self.ctx)
      self.ctx = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  # ---------------------------------------------------------------- marshalling
  def _use_stream(self):
    L.check(self.lib.cg_set_stream(self.ctx, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

  def to_device(self, x, shape=None):
    """numpy / torch / DLPack-able -> contiguous fp32 CUDA tensor."""
    if x is None:
      return None
    if not isinstance(x, torch.Tensor):
      if hasattr(x, '__dlpack__') and not isinstance(x, np.ndarray):
        x = torch.from_dlpack(x)
      else:
        x = torch.as_tensor(np.asarray(x))
    x = x.to(device=self.device, dtype=torch.float32).contiguous()
    if shape is not None and tuple(x.shape) != tuple(shape):
      raise ValueError('expected shape %s, got %s' % (tuple(shape), tuple(x.shape)))
    return x

  @staticmethod
  def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

  @staticmethod
  def _shifts(shifts, n):
    if shifts is None:
      return None
    a = np.ascontiguousarray(np.asarray(shifts, dtype=np.int32).reshape(-1))
    if a.size != n:
      raise ValueError('expected %d phase-shuffle shifts, got %d' % (n, a.size))
    return a.ctypes.data_as(C.POINTER(C.c_int32)), a

  def signal_shape(self, batch):
    return (batch, self.cfg.seq_len, self.cfg.channels)

  # ---------------------------------------------------------------- parameters
  def tensor_infos(self, which):
    out = []
    for i in range(self.lib.cg_num_tensors(self.ctx, which)):
      shape = (C.c_int64 * 4)()
      ndim, off = C.c_int(), C.c_int64()
      L.check(self.lib.cg_tensor_info(self.ctx, which, i, shape, C.byref(ndim), C.byref(off)))
      out.append((tuple(int(shape[j]) for j in range(ndim.value)), int(off.value)))
    return out

  def num_params(self, which):
    return int(self.lib.cg_num_params(self.ctx, which))

  def _split(self, which, flat):
    return [flat[o:o + int(np.prod(s))].reshape(s).copy() for s, o in self.tensor_infos(which)]

  def _join(self, which, arrays):
    infos = self.tensor_infos(which)
    if len(arrays) != len(infos):
      raise ValueError('expected %d arrays, got %d' % (len(infos), len(arrays)))
    flat = np.empty(self.num_params(which), np.float32)
    for a, (s, o) in zip(arrays, infos):
      a = np.asarray(a, dtype=np.float32)
      if tuple(a.shape) != tuple(s):
        raise ValueError('weight shape mismatch: expected %s, got %s' % (s, a.shape))
      flat[o:o + a.size] = a.reshape(-1)
    return flat

  def get_weights(self, which):
    self._use_stream()
    flat = np.empty(self.num_params(which), np.float32)
    L.check(self.lib.cg_get_weights(self.ctx, which, flat.ctypes.data_as(C.c_void_p)))
    return self._split(which, flat)

  def set_weights(self, which, arrays):
    self._use_stream()
    flat = self._join(which, arrays)
    L.check(self.lib.cg_set_weights(self.ctx, which, flat.ctypes.data_as(C.c_void_p)))

  def get_grads(self, which):
    self._use_stream()
    flat = np.empty(self.num_params(which), np.float32)
    L.check(self.lib.cg_get_grads(self.ctx, which, flat.ctypes.data_as(C.c_void_p)))
    return self._split(which, flat)

  def set_grad_buffer(self, which, tensor):
    """Accumulate gradients into `tensor` (flat fp32 CUDA, e.g. symmetric memory mapped by every peer); None restores."""
    if tensor is not None and (tensor.dtype != torch.float32 or tensor.numel() < self.num_params(which)):
      raise ValueError('gradient buffer must be float32 with at least num_params elements')
    L.check(self.lib.cg_set_grad_buffer(self.ctx, which, self._ptr(tensor)))
    self._grad_views.clear()
    self._ext_grad = getattr(self, '_ext_grad', {})
    self._ext_grad[which] = tensor

  def reduce_peer_grads(self, which, peer_ptrs, stream):
    """sum over ranks of the peers' gradient buffers -> the library's reduced buffer, on `stream` (cg_reduce_peer_grads)."""
    arr = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(int(p)) for p in peer_ptrs])
    L.check(self.lib.cg_reduce_peer_grads(self.ctx, which, arr, len(peer_ptrs), C.c_void_p(stream.cuda_stream)))

  def set_reduced_buffer(self, which, tensor):
    L.check(self.lib.cg_set_reduced_buffer(self.ctx, which, self._ptr(tensor)))
    self._ext_red = getattr(self, '_ext_red', {})
    self._ext_red[which] = tensor

  def peer_reduce_scatter(self, which, grad_peer_ptrs, rank, stream):
    arr = (C.c_void_p * len(grad_peer_ptrs))(*[C.c_void_p(int(p)) for p in grad_peer_ptrs])
    L.check(self.lib.cg_peer_reduce_scatter(self.ctx, which, arr, len(grad_peer_ptrs), int(rank), C.c_void_p(stream.cuda_stream)))

  def peer_all_gather(self, which, reduced_peer_ptrs, rank, stream):
    arr = (C.c_void_p * len(reduced_peer_ptrs))(*[C.c_void_p(int(p)) for p in reduced_peer_ptrs])
    L.check(self.lib.cg_peer_all_gather(self.ctx, which, arr, len(reduced_peer_ptrs), int(rank), C.c_void_p(stream.cuda_stream)))

  def apply_update_reduced(self, which):
    self._use_stream()
    L.check(self.lib.cg_apply_update_reduced(self.ctx, which))

  def reduced_grad_tensor(self, which):
    ptr = self.lib.cg_reduced_grad_ptr(self.ctx, which)
    return torch.as_tensor(_DevArray(ptr, self.num_params(which)), device=self.device)

  def set_grads(self, which, arrays):
    """Overwrite the flat gradient buffer (parity test of apply_update alone)."""
    flat = self._join(which, arrays)
    L.check(self.lib.cg_set_grads(self.ctx, which, flat.ctypes.data_as(C.c_void_p)))

  def skipped_updates(self, which):
    return int(self.lib.cg_skipped_updates(self.ctx, which))

  def grad_tensor(self, which):
    """Flat fp32 CUDA tensor aliasing the library's gradient buffer (for the NCCL all-reduce)."""
    if which not in self._grad_views:
      ptr = self.lib.cg_grad_ptr(self.ctx, which)
      self._grad_views[which] = torch.as_tensor(_DevArray(ptr, self.num_params(which)), device=self.device)
    return self._grad_views[which]

  def grad_buckets(self, which):
    """[(flat slice of the gradient tensor, bucket index)] in the order the backward pass completes them."""
    if ('b', which) not in self._grad_views:
      g, out = self.grad_tensor(which), []
      for b in range(self.lib.cg_num_buckets(self.ctx, which)):
        off, cnt = C.c_int64(), C.c_int64()
        L.check(self.lib.cg_bucket_info(self.ctx, which, b, C.byref(off), C.byref(cnt)))
        out.append((g[off.value:off.value + cnt.value], b))
      self._grad_views[('b', which)] = out
    return self._grad_views[('b', which)]

  def num_buckets(self, which):
    return int(self.lib.cg_num_buckets(self.ctx, which))

  def stream_wait_bucket(self, which, bucket, stream):
    L.check(self.lib.cg_stream_wait_bucket(self.ctx, which, bucket, C.c_void_p(stream.cuda_stream)))

  def init_weights(self, seed):
    self._use_stream()
    L.check(self.lib.cg_init_weights(self.ctx, C.c_uint64(int(seed))))

  def seed(self, seed):
    L.check(self.lib.cg_seed(self.ctx, C.c_uint64(int(seed))))

  def get_opt_state(self, which):
    n = self.num_params(which)
    m, v, step = np.empty(n, np.float32), np.empty(n, np.float32), C.c_int64()
    L.check(self.lib.cg_get_opt_state(self.ctx, which, m.ctypes.data_as(C.c_void_p),
                                      v.ctypes.data_as(C.c_void_p), C.byref(step)))
    return m, v, int(step.value)

  def set_opt_state(self, which, m=None, v=None, step=0):
    mp = np.ascontiguousarray(m, np.float32).ctypes.data_as(C.c_void_p) if m is not None else C.c_void_p(0)
    vp = np.ascontiguousarray(v, np.float32).ctypes.data_as(C.c_void_p) if v is not None else C.c_void_p(0)
    L.check(self.lib.cg_set_opt_state(self.ctx, which, mp, vp, C.c_int64(int(step))))

  def get_step(self, which):
    step = C.c_int64()
    L.check(self.lib.cg_get_opt_state(self.ctx, which, C.c_void_p(0), C.c_void_p(0), C.byref(step)))
    return int(step.value)

  def set_step(self, which, step):
    self.set_opt_state(which, None, None, step)

  # ---------------------------------------------------------------- hot path
  def prefetch_generator(self, real, noise=None, alpha=None, for_generator_step=False, want_fake32=True):
    """Generator part of the NEXT sub-step, enqueued now (cg_prefetch_generator)."""
    self._use_stream()
    real = self.to_device(real)
    B = real.shape[0]
    noise = self.to_device(noise, (B, self.cfg.noise_dim)) if noise is not None else None
    alpha = self.to_device(alpha).reshape(-1) if alpha is not None else None
    self._pref_keep = (real, noise, alpha)     # the library reads them when the kernels run
    L.check(self.lib.cg_prefetch_generator(self.ctx, self._ptr(real), B, self._ptr(noise), self._ptr(alpha),
                                           int(bool(for_generator_step)), 0 if want_fake32 else L.FLAG_NO_FAKE32))

  def critic_step(self, real, noise=None, alpha=None, shifts=None, update=True, sync=True, same_real=False,
                  want_fake32=True, gen_prefetched=False):
    self._use_stream()
    real = self.to_device(real)
    B = real.shape[0]
    noise = self.to_device(noise, (B, self.cfg.noise_dim)) if noise is not None else None
    alpha = self.to_device(alpha).reshape(-1) if alpha is not None else None
    sh = self._shifts(shifts, 12)
    flags = (0 if update else L.FLAG_NO_UPDATE) | (0 if sync else L.FLAG_NO_SYNC) | (L.FLAG_SAME_REAL if same_real else 0)
    flags |= 0 if want_fake32 else L.FLAG_NO_FAKE32
    flags |= L.FLAG_GEN_PREFETCHED if gen_prefetched else 0
    L.check(self.lib.cg_critic_step(self.ctx, self._ptr(real), B, self._ptr(noise), self._ptr(alpha),
                                    sh[0] if sh else None, flags, self._scal))
    return np.array(self._scal[:], np.float32) if sync else None

  def generator_step(self, real, noise=None, shifts=None, update=True, sync=True, gen_prefetched=False):
    self._use_stream()
    real = self.to_device(real)
    B = real.shape[0]
    noise = self.to_device(noise, (B, self.cfg.noise_dim)) if noise is not None else None
    sh = self._shifts(shifts, 4)
    flags = (0 if update else L.FLAG_NO_UPDATE) | (0 if sync else L.FLAG_NO_SYNC)
    flags |= L.FLAG_GEN_PREFETCHED if gen_prefetched else 0
    L.check(self.lib.cg_generator_step(self.ctx, self._ptr(real), B, self._ptr(noise),
                                       sh[0] if sh else None, flags, self._scal))
    return np.array(self._scal[:], np.float32) if sync else None

  def apply_update(self, which):
    self._use_stream()
    L.check(self.lib.cg_apply_update(self.ctx, which))

  def train_step(self, real, noises=None, alphas=None, shifts=None):
    self._use_stream()
    real = self.to_device(real)
    B, nc = real.shape[0], self.cfg.n_critic
    noises = self.to_device(noises, (nc + 1, B, self.cfg.noise_dim)) if noises is not None else None
    alphas = self.to_device(alphas, (nc, B)) if alphas is not None else None
    sh = self._shifts(shifts, 12 * nc + 4)
    L.check(self.lib.cg_train_step(self.ctx, self._ptr(real), B, self._ptr(noises), self._ptr(alphas),
                                   sh[0] if sh else None, self._scal))
    return np.array(self._scal[:], np.float32)

  def validate(self, real, noise=None, alpha=None, shifts=None):
    self._use_stream()
    real = self.to_device(real)
    B = real.shape[0]
    noise = self.to_device(noise, (B, self.cfg.noise_dim)) if noise is not None else None
    alpha = self.to_device(alpha).reshape(-1) if alpha is not None else None
    sh = self._shifts(shifts, 12)
    fake = torch.empty(self.signal_shape(B), device=self.device, dtype=torch.float32)
    L.check(self.lib.cg_validate(self.ctx, self._ptr(real), B, self._ptr(noise), self._ptr(alpha),
                                 sh[0] if sh else None, self._ptr(fake), self._scal))
    return fake, np.array(self._scal[:], np.float32)

  def gather_rows(self, src, idx, out=None):
    """out[i] = src[idx[i]] on the device (cg_gather_rows); src (N, ...) fp32 CUDA, idx int64 CUDA."""
    self._use_stream()
    n, row = int(idx.numel()), int(src[0].numel())
    if out is None:
      out = torch.empty((n,) + tuple(src.shape[1:]), device=self.device, dtype=torch.float32)
    L.check(self.lib.cg_gather_rows(self.ctx, self._ptr(src), int(src.shape[0]), C.c_void_p(idx.data_ptr()), n, row,
                                    self._ptr(out)))
    return out

  def metrics(self, real, fake):
    """gan.py:32-41 on caller tensors -> [min, max, mean, std] errors."""
    self._use_stream()
    real, fake = self.to_device(real), self.to_device(fake)
    if real.shape != fake.shape or tuple(real.shape[1:]) != self.signal_shape(1)[1:]:
      raise ValueError('metrics: expected two (batch, %d, %d) tensors' % self.signal_shape(1)[1:])
    out = (C.c_float * 4)()
    L.check(self.lib.cg_metrics(self.ctx, self._ptr(real), self._ptr(fake), real.shape[0], out))
    return [float(x) for x in out]

  def generate(self, noise, denorm=False):
    self._use_stream()
    noise = self.to_device(noise)
    B = noise.shape[0]
    out = torch.empty(self.signal_shape(B), device=self.device, dtype=torch.float32)
    L.check(self.lib.cg_generate(self.ctx, self._ptr(noise), B, int(bool(denorm)), self._ptr(out)))
    return out

  def critic_forward(self, x, shifts):
    self._use_stream()
    x = self.to_device(x)
    B = x.shape[0]
    sh = self._shifts(shifts, 4)
    out = torch.empty((B,), device=self.device, dtype=torch.float32)
    L.check(self.lib.cg_debug_critic_forward(self.ctx, self._ptr(x), B, sh[0], self._ptr(out)))
    return out.reshape(B, 1)

  def gp_debug(self, xhat, shifts):
    """returns (dD/dxhat (B,L,C), squared norms (B,))"""
    self._use_stream()
    x = self.to_device(xhat)
    B = x.shape[0]
    sh = self._shifts(shifts, 4)
    g = torch.empty(self.signal_shape(B), device=self.device, dtype=torch.float32)
    n2 = torch.empty((B,), device=self.device, dtype=torch.float32)
    L.check(self.lib.cg_debug_gp(self.ctx, self._ptr(x), B, sh[0], self._ptr(g), self._ptr(n2)))
    return g, n2

  def gp_gradient(self, xhat, shifts, sync=True):
    """Gradient penalty at xhat and gp_lambda * dGP/dW (cg_gp_gradient); returns the GP value (None without sync)."""
    self._use_stream()
    x = self.to_device(xhat)
    sh = self._shifts(shifts, 4)
    L.check(self.lib.cg_gp_gradient(self.ctx, self._ptr(x), x.shape[0], sh[0], 0 if sync else L.FLAG_NO_SYNC, self._scal))
    return float(self._scal[L.S_GP]) if sync else None

  def debug_layer(self, which, layer, pass_, x=None, dy=None):
    """One conv layer in isolation (cg_debug_layer): fp32 tensors in / out."""
    self._use_stream()
    x, dy = self.to_device(x), self.to_device(dy)
    B = (x if x is not None else dy).shape[0]
    infos = self.tensor_infos(which)
    if which == L.DISCRIMINATOR:
      K, cin, cout = infos[2 * (layer - 1)][0]
      lin = self.cfg.seq_len >> (layer - 1)
      lout = lin // 2
      kshape = (K, cin, cout)
    else:
      idx = 2 + (layer - 1) * (4 if self.cfg.layer_norm else 2)
      K, _, cout, cin = infos[idx][0]
      lin = (self.cfg.seq_len // 32) << (layer - 1)
      lout = lin * 2
      kshape = (K, 1, cout, cin)
    shape = {0: (B, lout, cout), 1: (B, lin, cin), 2: kshape}[pass_]
    out = torch.empty(shape, device=self.device, dtype=torch.float32)
    L.check(self.lib.cg_debug_layer(self.ctx, which, layer, pass_, self._ptr(x), self._ptr(dy), B, self._ptr(out)))
    return out

  def phase_shuffle(self, x, shift):
    self._use_stream()
    x = self.to_device(x)
    B, w, ch = x.shape
    out = torch.empty_like(x)
    L.check(self.lib.cg_debug_phase_shuffle(self.ctx, self._ptr(x), B, w, ch, int(shift), self._ptr(out)))
    return out

  def debug_read(self, buffer, layer, batch):
    """Internal activation buffer of the last step as an unpadded fp32 CUDA tensor (batch, rows, channels)."""
    self._use_stream()
    rows, ch = C.c_int64(), C.c_int64()
    L.check(self.lib.cg_debug_buffer_shape(self.ctx, buffer, layer, C.byref(rows), C.byref(ch)))
    out = torch.empty((batch, rows.value, ch.value), device=self.device, dtype=torch.float32)
    L.check(self.lib.cg_debug_read(self.ctx, buffer, layer, batch, self._ptr(out)))
    return out

  def last_draws(self, n_noise=0, n_alpha=0, n_shifts=0):
    """(noise, alpha, shifts) of the last step function, injected or library-drawn."""
    self._use_stream()
    noise = torch.empty((n_noise,), device=self.device, dtype=torch.float32) if n_noise else None
    alpha = torch.empty((n_alpha,), device=self.device, dtype=torch.float32) if n_alpha else None
    sh = np.zeros(max(n_shifts, 1), np.int32)
    L.check(self.lib.cg_debug_last_draws(self.ctx, self._ptr(noise), n_noise, self._ptr(alpha), n_alpha,
                                         sh.ctypes.data_as(C.POINTER(C.c_int32)) if n_shifts else None, n_shifts))
    return noise, alpha, sh[:n_shifts]

  def dgrad_ps(self, layer, dy, h, group_b, shifts):
    """DA[l] -> DA[l-1] in isolation: data gradient + PhaseShuffle adjoint + LeakyReLU slope (cg_debug_dgrad_ps)."""
    self._use_stream()
    dy, h = self.to_device(dy), self.to_device(h)
    sh = np.ascontiguousarray(np.asarray(shifts, np.int32).reshape(-1))
    out = torch.empty_like(h)
    L.check(self.lib.cg_debug_dgrad_ps(self.ctx, layer, self._ptr(dy), self._ptr(h), dy.shape[0], int(group_b),
                                       sh.ctypes.data_as(C.POINTER(C.c_int32)), self._ptr(out)))
    return out

  def fake(self, batch):
    """fp32 CUDA tensor aliasing the generator output of the last step."""
    ptr = self.lib.cg_fake_ptr(self.ctx)
    n = batch * self.cfg.seq_len * self.cfg.channels
    return torch.as_tensor(_DevArray(ptr, n), device=self.device).reshape(self.signal_shape(batch))

  def scores(self, n):
    ptr = self.lib.cg_scores_ptr(self.ctx)
    return torch.as_tensor(_DevArray(ptr, n), device=self.device)

  def scalars_tensor(self):
    """CUDA view of the scalars the last step wrote (slot 0), valid on the engine stream."""
    ptr = self.lib.cg_scalars_ptr(self.ctx)
    return torch.as_tensor(_DevArray(ptr, L.NUM_SCALARS), device=self.device)

  def launch_count(self):
    return int(self.lib.cg_launch_count(self.ctx))

  def tc_launch_count(self):
    return int(self.lib.cg_tc_launch_count(self.ctx))

  def device_bytes(self):
    return int(self.lib.cg_device_bytes(self.ctx))

  def synchronize(self):
    L.check(self.lib.cg_synchronize(self.ctx))

  def profile(self, enable):
    L.check(self.lib.cg_profile(self.ctx, int(bool(enable))))

  def profile_report(self):
    out = (C.c_double * 12)()
    L.check(self.lib.cg_profile_report(self.ctx, out))
    return {'gemm': {'ms': out[0], 'flops': out[1], 'launches': int(out[2])},
            'wgrad': {'ms': out[3], 'flops': out[4], 'launches': int(out[5])},
            'head': {'ms': out[6], 'flops': out[7], 'launches': int(out[8])}}

  def profile_table(self):
    """{kernel: {bound, launches, ms, flops, bytes}} of everything timed since profile(True) (cg_profile_report_text)."""
    buf = C.create_string_buffer(1 << 16)
    L.check(self.lib.cg_profile_report_text(self.ctx, buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
      name, bound, n, ms, fl, by = line.split('\t')
      out[name] = {'bound': bound, 'launches': int(n), 'ms': float(ms), 'flops': float(fl), 'bytes': float(by)}
    return out

  def bench_layer(self, which, layer, pass_, batch, iters=10):
    self._use_stream()
    ms, fl = C.c_float(), C.c_double()
    L.check(self.lib.cg_bench_layer(self.ctx, which, layer, pass_, batch, iters, C.byref(ms), C.byref(fl)))
    return float(ms.value), float(fl.value)
