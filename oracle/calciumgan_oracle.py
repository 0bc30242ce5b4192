"""CPU oracle for the CalciumGAN WGAN-GP training step.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED AT THE TENSORFLOW LEVEL: the reference (bryanlimy/CalciumGAN) ships no tests,
golden vectors or fixtures, and its arithmetic lives in un-vendored TensorFlow 2.3.1
(setup.sh:26,29), which cannot be installed in this image.  What IS pinned: the reference's own
Python files (model builders, PhaseShuffle, WGAN-GP losses / tapes / update order, metrics) are
executed unmodified from /root/reference over a torch-backed stand-in for the TF / Keras calls
they make (oracle/tf_shim, oracle/reference_runner.py), and this restatement reproduces their
train step, validate and generate to 1e-9 in float64 (tests/test_reference_shim.py; committed
fixture tests/golden/reference_step.npz + its generating script).  The semantics of the TF
kernels themselves (Keras defaults below) remain restated, here and -- independently -- in the
stand-in.  Self-consistency checks on top: naive-loop convolutions vs torch, autograd vs the
hand-derived 4-pass gradient penalty, finite differences, literal pad+slice PhaseShuffle vs
the closed form.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  The product path (``calciumgan_b200``) never does.

Reference files followed (paths relative to /root/reference):
  gan/models/calciumgan.py:15-19    noise width            -> calculate_noise_shape
  gan/models/calciumgan.py:22-103   generator              -> generator_forward
  gan/models/calciumgan.py:106-138  PhaseShuffle           -> phase_shuffle / phase_shuffle_literal
  gan/models/calciumgan.py:141-192  critic                 -> discriminator_forward
  gan/models/utils.py:6-8           LeakyReLU (alpha 0.3)  -> leaky_relu
  gan/models/utils.py:65-94         Conv1DTranspose        -> conv1d_transpose_same
  gan/algorithms/wgan_gp.py:19-95   WGAN-GP losses / steps -> critic_step, generator_step, train_step
  gan/algorithms/gan.py:29-41       noise, metrics         -> signals_metrics
  gan/algorithms/optimizer.py:7-34  Adam(lr) update        -> adam_update
  gan/utils/signals_metrics.py:9-28 min/max/mean/std error -> signals_metrics
  gan/utils/utils.py:30-32          denormalize            -> denormalize

TF 2.3.1 / Keras defaults this restatement relies on (the main parity risk):
  LeakyReLU alpha = 0.3, gradient at 0 is alpha;  LayerNormalization axis=-1, eps=1e-3,
  biased variance;  SAME padding: total = K - s (L % s == 0), left = floor(total / 2);
  Conv2DTranspose kernel (kh, kw, out, in), output length = in * stride;
  Flatten is row-major over (time, channel);  tape.gradient of a non-scalar = grad of its sum;
  tf.norm has no epsilon;  Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t), w -= lr_t*m/(sqrt(v)+1e-7);
  reduce_std = population std.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LEAKY_ALPHA = 0.3   # tf.keras.layers.LeakyReLU() default (models/utils.py:7)
LN_EPS = 1e-3       # tf.keras.layers.LayerNormalization() default (calciumgan.py:45)
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-7   # tf.keras.optimizers.Adam defaults (optimizer.py:9)
NUM_LAYERS = 5


@dataclass
class HParams:
  """The hparams fields the hot path reads (README.md:92 paper values as defaults)."""
  signal_shape: Tuple[int, int] = (2048, 102)
  noise_dim: int = 32
  num_units: int = 64
  kernel_size: int = 24
  strides: int = 2
  m: int = 10
  layer_norm: bool = True
  batch_norm: bool = False
  activation: str = 'leakyrelu'
  normalize: bool = True
  signals_min: float = 0.0
  signals_max: float = 1.0
  gradient_penalty: float = 10.0
  n_critic: int = 5
  learning_rate: float = 1e-4
  conv2d: bool = False

  @property
  def num_channels(self) -> int:
    return self.signal_shape[1]

  @property
  def noise_shape(self) -> Tuple[int]:
    return (self.noise_dim,)


# ----------------------------------------------------------------------------- shapes

def calculate_noise_shape(output_shape, noise_dim, num_convolutions, strides):
  """calciumgan.py:15-19."""
  w = output_shape[0] / (strides**num_convolutions)
  if not float(w).is_integer():
    raise ValueError('Conv1D: w {} is not an integer.'.format(w))
  return (int(w), noise_dim)


def generator_channels(hp: HParams) -> List[int]:
  """Cin of convT1 .. Cout of convT5 (calciumgan.py:37-89)."""
  nu = hp.num_units
  return [hp.noise_dim, nu * 5, nu * 4, nu * 3, nu * 2, hp.num_channels]


def discriminator_channels(hp: HParams) -> List[int]:
  """Cin of conv1 .. Cout of conv5 (calciumgan.py:145-185)."""
  nu = hp.num_units
  return [hp.num_channels, nu, nu * 2, nu * 3, nu * 4, nu * 5]


def weight_shapes(hp: HParams):
  """get_weights() order/layout of both Keras models (SURVEY §8a last row)."""
  K = hp.kernel_size
  w, nd = calculate_noise_shape(hp.signal_shape, hp.noise_dim, NUM_LAYERS, hp.strides)
  gc, dc = generator_channels(hp), discriminator_channels(hp)
  gen = [(nd, w * nd), (w * nd,)]
  for i in range(NUM_LAYERS):
    gen += [(K, 1, gc[i + 1], gc[i]), (gc[i + 1],)]
    if hp.layer_norm:
      gen += [(gc[i + 1],), (gc[i + 1],)]
  gen += [(hp.num_channels, hp.num_channels), (hp.num_channels,)]
  dis = []
  for i in range(NUM_LAYERS):
    dis += [(K, dc[i], dc[i + 1]), (dc[i + 1],)]
  w5 = hp.signal_shape[0] // (hp.strides**NUM_LAYERS)
  dis += [(w5 * dc[5], 1), (1,)]
  return gen, dis


def _glorot(rng, shape, fan_in, fan_out):
  limit = math.sqrt(6.0 / (fan_in + fan_out))
  return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def init_weights(hp: HParams, seed: int = 1234):
  """Keras default init: glorot-uniform kernels, zero biases, LN gamma=1 / beta=0."""
  rng = np.random.RandomState(seed)
  gs, ds = weight_shapes(hp)

  def fill(shapes, is_gen):
    out, i = [], 0
    while i < len(shapes):
      s = shapes[i]
      if len(s) == 2:      # dense (in, out)
        out.append(_glorot(rng, s, s[0], s[1]))
      elif len(s) == 3:    # conv1d (K, Cin, Cout)
        out.append(_glorot(rng, s, s[0] * s[1], s[0] * s[2]))
      elif len(s) == 4:    # conv2d-transpose (K, 1, Cout, Cin): fan_in = K*Cin, fan_out = K*Cout
        out.append(_glorot(rng, s, s[0] * s[3], s[0] * s[2]))
      out.append(np.zeros(shapes[i + 1], np.float32))
      i += 2
      if is_gen and hp.layer_norm and len(s) == 4:
        out.append(np.ones(shapes[i], np.float32))
        out.append(np.zeros(shapes[i + 1], np.float32))
        i += 2
    return out

  return fill(gs, True), fill(ds, False)


def randomize_weights(weights: Sequence[np.ndarray], seed: int, scale: float = 0.05):
  """Perturb biases / LN params so parity tests exercise them (they init to 0 / 1)."""
  rng = np.random.RandomState(seed)
  out = []
  for w in weights:
    if w.ndim == 1:
      out.append((w + scale * rng.standard_normal(w.shape)).astype(np.float32))
    else:
      out.append(w.copy())
  return out


# ----------------------------------------------------------------------------- layers

def leaky_relu(x, slope=None):
  """LeakyReLU; with ``slope`` (a tensor of 1 / LEAKY_ALPHA per element) the branch of every element is IMPOSED
  instead of read from the sign of x. That turns the network into a smooth function of its rounding errors
  (LeakyReLU's slope is its only discontinuity), which is how the parity tests separate "an activation rounded
  across zero" from "a kernel bug": the CUDA path exports its own branch decisions (sign of the stored activation)
  and the oracle is evaluated on the same branches."""
  if slope is not None:
    return x * slope
  return torch.where(x > 0, x, LEAKY_ALPHA * x)


def slopes_from_activation(h):
  """Branch decisions of a stored LeakyReLU output (sign(h) == sign(pre-activation); 0 counts as negative, as in
  the engine's lrelu_slope)."""
  return torch.where(h > 0, torch.ones_like(h), torch.full_like(h, LEAKY_ALPHA))


def same_pad_left(K: int, s: int) -> int:
  return max(K - s, 0) // 2


def conv1d_same(x, kernel, bias, stride=2):
  """Keras Conv1D(padding='same', strides=2) on NWC input (calciumgan.py:145-149).
  kernel (K, Cin, Cout); y[b,o,co] = sum_k,ci x[b, s*o + k - padL, ci] * W[k,ci,co] + b."""
  K = kernel.shape[0]
  total = max(K - stride, 0)
  left = total // 2
  xt = F.pad(x.permute(0, 2, 1), (left, total - left))
  y = F.conv1d(xt, kernel.permute(2, 1, 0), bias, stride=stride)
  return y.permute(0, 2, 1)


def conv1d_transpose_same(x, kernel, bias, stride=2):
  """models/utils.py:65-94 (Conv2DTranspose (K,1)/(s,1) 'same' on an expanded axis).
  kernel (K, 1, Cout, Cin); y[b, s*i + k - padL, co] += x[b,i,ci] * W[k,0,co,ci]; Lout = s*Lin."""
  K = kernel.shape[0]
  left = same_pad_left(K, stride)
  Lin = x.shape[1]
  w = kernel[:, 0].permute(2, 1, 0)   # (Cin, Cout, K)
  full = F.conv_transpose1d(x.permute(0, 2, 1), w, None, stride=stride)
  need = left + stride * Lin
  if full.shape[-1] < need:
    full = F.pad(full, (0, need - full.shape[-1]))
  y = full[..., left:need].permute(0, 2, 1)
  return y + bias if bias is not None else y


def layer_norm(x, gamma, beta):
  mu = x.mean(dim=-1, keepdim=True)
  var = ((x - mu)**2).mean(dim=-1, keepdim=True)
  return (x - mu) / torch.sqrt(var + LN_EPS) * gamma + beta


def phase_shuffle_index(w: int, shift: int) -> np.ndarray:
  """Closed form of calciumgan.py:117-138: out[:, t] = x[:, idx[t]]. int32, must be bit-exact."""
  t = np.arange(w, dtype=np.int64) + int(shift)
  j = np.abs(t)
  j = np.where(j > w - 1, 2 * (w - 1) - j, j)
  return j.astype(np.int32)


def phase_shuffle(x, shift: int):
  idx = torch.from_numpy(phase_shuffle_index(x.shape[1], shift).astype(np.int64)).to(x.device)
  return x.index_select(1, idx)


def phase_shuffle_literal(x: np.ndarray, shift: int) -> np.ndarray:
  """Literal transcription of calciumgan.py:126-137 with numpy reflect padding."""
  w = x.shape[1]
  if shift > 0:
    paddings, start, end = [(0, 0), (0, shift), (0, 0)], shift, w + shift
  else:
    paddings, start, end = [(0, 0), (abs(shift), 0), (0, 0)], 0, w
  return np.pad(x, paddings, mode='reflect')[:, start:end, :]


# ----------------------------------------------------------------------------- models

_DEVICE = None   # tests may evaluate the oracle with torch float64 on the CUDA device for batch-128 cases (set_device)


def set_device(device):
  """Where the oracle's tensors live. Default (None) = CPU. The arithmetic is the same torch code either way; the
  full-batch parity test moves it to the GPU because a batch-128 float64 double backward takes minutes on the host."""
  global _DEVICE
  _DEVICE = device


def _t(a, dtype):
  if isinstance(a, torch.Tensor):
    return a.to(device=_DEVICE, dtype=dtype) if (_DEVICE is not None or a.dtype != dtype) else a
  return torch.as_tensor(np.asarray(a), dtype=dtype, device=_DEVICE)


def generator_forward(gw, noise, hp: HParams, taps: Optional[dict] = None, slopes=None):
  """calciumgan.py:22-103. gw = list of tensors in get_weights() order.
  slopes: optional imposed LeakyReLU branches [dense, convT1 .. convT5] (see leaky_relu)."""
  if hp.batch_norm:
    raise NotImplementedError('batch_norm is out of scope (SURVEY §3.6-6)')
  w, nd = calculate_noise_shape(hp.signal_shape, hp.noise_dim, NUM_LAYERS, hp.strides)
  i = 0
  x = noise @ gw[0] + gw[1]
  x = leaky_relu(x, None if slopes is None else slopes[0].reshape(x.shape))
  if taps is not None:
    taps['g0'] = x
  i = 2
  x = x.reshape(x.shape[0], w, nd)
  for l in range(NUM_LAYERS):
    x = conv1d_transpose_same(x, gw[i], gw[i + 1], hp.strides)
    i += 2
    if taps is not None:
      taps['convt%d' % (l + 1)] = x
    if hp.layer_norm:
      x = layer_norm(x, gw[i], gw[i + 1])
      i += 2
    x = leaky_relu(x, None if slopes is None else slopes[l + 1])
    if taps is not None:
      taps['g%d' % (l + 1)] = x
  x = x @ gw[i] + gw[i + 1]
  return torch.sigmoid(x) if hp.normalize else x


def discriminator_forward(dw, x, shifts: Sequence[int], hp: HParams, taps: Optional[dict] = None, slopes=None):
  """calciumgan.py:141-192; ``shifts`` = the 4 PhaseShuffle draws of this call.
  slopes: optional imposed LeakyReLU branches [conv1 .. conv5] (see leaky_relu)."""
  for l in range(NUM_LAYERS):
    x = leaky_relu(conv1d_same(x, dw[2 * l], dw[2 * l + 1], hp.strides), None if slopes is None else slopes[l])
    if taps is not None:
      taps['h%d' % (l + 1)] = x
    if l < NUM_LAYERS - 1:
      x = phase_shuffle(x, int(shifts[l]))
  x = x.reshape(x.shape[0], -1)
  return x @ dw[10] + dw[11]


# ----------------------------------------------------------------------------- losses / steps

def denormalize(x, x_min, x_max):
  return x * (x_max - x_min) + x_min


def signals_metrics(real, fake, hp: HParams) -> Dict[str, float]:
  """gan.py:32-41 + signals_metrics.py:9-28."""
  if hp.normalize:
    real = denormalize(real, hp.signals_min, hp.signals_max)
    fake = denormalize(fake, hp.signals_min, hp.signals_max)

  def std(x):
    return torch.sqrt(((x - x.mean(-1, keepdim=True))**2).mean(-1))

  return {
      'signals_metrics/min': float(((real.min(-1).values - fake.min(-1).values)**2).mean()),
      'signals_metrics/max': float(((real.max(-1).values - fake.max(-1).values)**2).mean()),
      'signals_metrics/mean': float(((real.mean(-1) - fake.mean(-1))**2).mean()),
      'signals_metrics/std': float(((std(real) - std(fake))**2).mean()),
  }


def gradient_penalty(dw, real, fake, alpha, shifts, hp, create_graph=True, slopes=None, taps=None):
  """wgan_gp.py:38-50. alpha (B,1,1)."""
  xhat = (alpha * real + (1 - alpha) * fake).detach().requires_grad_(True)
  out = discriminator_forward(dw, xhat, shifts, hp, taps, slopes)
  (g,) = torch.autograd.grad(out.sum(), xhat, create_graph=create_graph)
  norm = torch.sqrt((g.reshape(g.shape[0], -1)**2).sum(dim=1))
  return ((norm - 1.0)**2).mean(), g, norm


def critic_step(gw, dw, real, noise, alpha, shifts, hp: HParams, dtype=torch.float64, slopes=None, fake=None):
  """wgan_gp.py:64-80 without the optimizer: returns losses + per-parameter gradients.
  shifts: (3, 4) ints = draws of D(real), D(fake), D(xhat) in that call order.
  slopes: optional {'real' | 'fake' | 'xhat': [5 tensors]} imposed LeakyReLU branches of the three critic calls
  (and 'gen': [6 tensors] for the generator forward); the free-running branch decisions are returned under 'acts'."""
  gw = [_t(a, dtype) for a in gw]
  dw = [_t(a, dtype).clone().requires_grad_(True) for a in dw]
  real, noise, alpha = _t(real, dtype), _t(noise, dtype), _t(alpha, dtype).reshape(-1, 1, 1)
  sl = slopes or {}
  taps = {'real': {}, 'fake': {}, 'xhat': {}}
  with torch.no_grad():
    fake = generator_forward(gw, noise, hp, slopes=sl.get('gen')) if fake is None else _t(fake, dtype)
  real_out = discriminator_forward(dw, real, shifts[0], hp, taps['real'], sl.get('real'))
  fake_out = discriminator_forward(dw, fake, shifts[1], hp, taps['fake'], sl.get('fake'))
  gp, g, norm = gradient_penalty(dw, real, fake, alpha, shifts[2], hp, slopes=sl.get('xhat'), taps=taps['xhat'])
  loss = -real_out.mean() + fake_out.mean() + hp.gradient_penalty * gp
  grads = torch.autograd.grad(loss, dw)
  return {
      'dis_loss': float(loss.detach()), 'gradient_penalty': float(gp.detach()),
      'fake': fake.detach(), 'real_out': real_out.detach(), 'fake_out': fake_out.detach(),
      'gp_grad': g.detach(), 'gp_norm': norm.detach(),
      'grads': [x.detach() for x in grads],
      'acts': {k: [v['h%d' % (l + 1)].detach() for l in range(NUM_LAYERS)] for k, v in taps.items()},
  }


def generator_step(gw, dw, real, noise, shifts, hp: HParams, dtype=torch.float64, slopes=None):
  """wgan_gp.py:22-36 without the optimizer. shifts: 4 ints (draws of D(fake)).
  slopes: optional {'gen': [6 tensors], 'fake': [5 tensors]} imposed LeakyReLU branches (see leaky_relu)."""
  gw = [_t(a, dtype).clone().requires_grad_(True) for a in gw]
  dw = [_t(a, dtype) for a in dw]
  noise = _t(noise, dtype)
  sl = slopes or {}
  gtaps, dtaps = {}, {}
  fake = generator_forward(gw, noise, hp, gtaps, sl.get('gen'))
  fake_out = discriminator_forward(dw, fake, shifts, hp, dtaps, sl.get('fake'))
  loss = -fake_out.mean()
  grads = torch.autograd.grad(loss, gw)
  out = {'gen_loss': float(loss.detach()), 'fake': fake.detach(), 'fake_out': fake_out.detach(),
         'grads': [x.detach() for x in grads],
         'acts': {'gen': [gtaps['g%d' % l].detach() for l in range(NUM_LAYERS + 1)],
                  'fake': [dtaps['h%d' % (l + 1)].detach() for l in range(NUM_LAYERS)]}}
  if real is not None:
    out['metrics'] = signals_metrics(_t(real, dtype), fake.detach(), hp)
  return out


def adam_update(w, m, v, g, t: int, lr: float, b1: float = ADAM_B1, b2: float = ADAM_B2):
  """Keras Adam dense update at iteration t (1-based, i.e. iterations+1) (optimizer.py:34)."""
  lr_t = lr * math.sqrt(1.0 - b2**t) / (1.0 - b1**t)
  m = b1 * m + (1 - b1) * g
  v = b2 * v + (1 - b2) * g * g
  w = w - lr_t * m / (torch.sqrt(v) + ADAM_EPS)
  return w, m, v


@dataclass
class TrainState:
  gen: List[torch.Tensor]
  dis: List[torch.Tensor]
  gen_m: List[torch.Tensor] = field(default_factory=list)
  gen_v: List[torch.Tensor] = field(default_factory=list)
  dis_m: List[torch.Tensor] = field(default_factory=list)
  dis_v: List[torch.Tensor] = field(default_factory=list)
  gen_steps: int = 0
  dis_steps: int = 0

  @staticmethod
  def create(gen_w, dis_w, dtype=torch.float64):
    g = [_t(a, dtype).clone() for a in gen_w]
    d = [_t(a, dtype).clone() for a in dis_w]
    z = lambda ws: [torch.zeros_like(a) for a in ws]
    return TrainState(g, d, z(g), z(g), z(d), z(d))


def train_step(state: TrainState, real, noises, alphas, shifts, hp: HParams, dtype=torch.float64):
  """wgan_gp.py:82-95: n_critic critic updates on the SAME real batch, then one generator update.
  noises: (n_critic+1, B, nd); alphas: (n_critic, B); shifts: (n_critic*12 + 4,) ints in draw order
  (per critic sub-step: 4 for D(real), 4 for D(fake), 4 for D(xhat); then 4 for the generator step)."""
  shifts = np.asarray(shifts).reshape(-1)
  dis_losses, gps = [], []
  for i in range(hp.n_critic):
    sh = shifts[12 * i:12 * i + 12].reshape(3, 4)
    r = critic_step(state.gen, state.dis, real, noises[i], alphas[i], sh, hp, dtype)
    state.dis_steps += 1
    for j, g in enumerate(r['grads']):
      state.dis[j], state.dis_m[j], state.dis_v[j] = adam_update(
          state.dis[j], state.dis_m[j], state.dis_v[j], g, state.dis_steps, hp.learning_rate)
    dis_losses.append(r['dis_loss'])
    gps.append(r['gradient_penalty'])
  sh = shifts[12 * hp.n_critic:12 * hp.n_critic + 4]
  r = generator_step(state.gen, state.dis, real, noises[hp.n_critic], sh, hp, dtype)
  state.gen_steps += 1
  for j, g in enumerate(r['grads']):
    state.gen[j], state.gen_m[j], state.gen_v[j] = adam_update(
        state.gen[j], state.gen_m[j], state.gen_v[j], g, state.gen_steps, hp.learning_rate)
  return r['gen_loss'], float(np.mean(dis_losses)), float(np.mean(gps)), r['metrics']


def validate_step(gw, dw, real, noise, alpha, shifts, hp: HParams, dtype=torch.float64):
  """gan.py:58-70,87-90 with WGAN-GP losses (wgan_gp.py:19-20,52-62), training=False.
  shifts: (3, 4): D(real), D(fake), D(xhat)."""
  gw = [_t(a, dtype) for a in gw]
  dw = [_t(a, dtype) for a in dw]
  real, noise, alpha = _t(real, dtype), _t(noise, dtype), _t(alpha, dtype).reshape(-1, 1, 1)
  with torch.no_grad():
    fake = generator_forward(gw, noise, hp)
    real_out = discriminator_forward(dw, real, shifts[0], hp)
    fake_out = discriminator_forward(dw, fake, shifts[1], hp)
  gp, _, _ = gradient_penalty(dw, real, fake, alpha, shifts[2], hp, create_graph=False)
  gen_loss = float(-fake_out.mean())
  dis_loss = float(-real_out.mean() + fake_out.mean() + hp.gradient_penalty * gp)
  return fake, gen_loss, dis_loss, float(gp), signals_metrics(real, fake, hp)


# ----------------------------------------------------------------------------- hand-derived GP

def _ps_transpose(dx, shift):
  """Adjoint of phase_shuffle: scatter-add along the same index map."""
  w = dx.shape[1]
  idx = torch.from_numpy(phase_shuffle_index(w, shift).astype(np.int64)).to(dx.device)
  out = torch.zeros_like(dx)
  return out.index_add(1, idx, dx)


def conv1d_same_dgrad(dy, kernel, Lin, stride=2):
  """dx of conv1d_same: dx[b,i,ci] = sum_{o,k: s*o+k-padL=i} dy[b,o,co] W[k,ci,co]."""
  K = kernel.shape[0]
  left = same_pad_left(K, stride)
  w = kernel.permute(2, 1, 0)   # (Cout, Cin, K) == conv_transpose weight (in=Cout, out=Cin, K)
  full = F.conv_transpose1d(dy.permute(0, 2, 1), w, None, stride=stride)
  need = left + Lin
  if full.shape[-1] < need:
    full = F.pad(full, (0, need - full.shape[-1]))
  return full[..., left:need].permute(0, 2, 1)


def conv1d_same_wgrad(x, dy, K, stride=2):
  """dW[k,ci,co] = sum_{b,o} x[b, s*o+k-padL, ci] dy[b,o,co]."""
  total = max(K - stride, 0)
  left = total // 2
  xp = F.pad(x, (0, 0, left, total - left))
  Lout = dy.shape[1]
  out = []
  for k in range(K):
    xs = xp[:, k:k + stride * Lout:stride, :]
    out.append(torch.einsum('boi,boc->ic', xs, dy))
  return torch.stack(out)


def gp_four_pass(dw, xhat, shifts, hp: HParams, dtype=torch.float64, slopes=None):
  """The 4-pass gradient-penalty gradient that never builds the second-order graph
  (SURVEY §8a): forward (masks) -> dgrad chain (g) -> linearised forward of u (no biases)
  -> wgrad(v_{l-1}, delta_l).  Returns gp, g, and dGP/dW for all 12 critic tensors
  (NOT multiplied by the penalty weight). slopes: optional imposed LeakyReLU branches [conv1 .. conv5]."""
  dw = [_t(a, dtype) for a in dw]
  x = _t(xhat, dtype)
  B = x.shape[0]
  K = hp.kernel_size
  masks, lens = [], []
  for l in range(NUM_LAYERS):
    lens.append(x.shape[1])
    a = conv1d_same(x, dw[2 * l], dw[2 * l + 1])
    masks.append(slopes_from_activation(a) if slopes is None else _t(slopes[l], dtype))
    x = a * masks[-1]
    if l < NUM_LAYERS - 1:
      x = phase_shuffle(x, int(shifts[l]))
  # pass 2: dgrad chain of sum_b D(xhat)_b
  deltas = [None] * NUM_LAYERS
  d = masks[4] * dw[10].reshape(1, masks[4].shape[1], masks[4].shape[2])
  deltas[4] = d
  for l in range(NUM_LAYERS - 2, -1, -1):
    dx = conv1d_same_dgrad(deltas[l + 1], dw[2 * (l + 1)], lens[l + 1])
    deltas[l] = masks[l] * _ps_transpose(dx, int(shifts[l]))
  g = conv1d_same_dgrad(deltas[0], dw[0], lens[0])
  n = torch.sqrt((g.reshape(B, -1)**2).sum(1))
  gp = ((n - 1.0)**2).mean()
  # pass 3: linearised forward of u = dGP/dg
  u = (2.0 / B) * ((n - 1.0) / n).reshape(B, 1, 1) * g
  grads = [torch.zeros_like(w) for w in dw]
  v = u
  for l in range(NUM_LAYERS):
    grads[2 * l] = conv1d_same_wgrad(v, deltas[l], K)          # pass 4
    c = masks[l] * conv1d_same(v, dw[2 * l], None)
    v = phase_shuffle(c, int(shifts[l])) if l < NUM_LAYERS - 1 else c
  grads[10] = v.sum(0).reshape(-1, 1)
  return gp, g, grads


# ----------------------------------------------------------------------------- naive definitional loops (small cases)

def naive_conv1d_same(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray, stride=2) -> np.ndarray:
  B, L, Cin = x.shape
  K, _, Cout = kernel.shape
  Lout = -(-L // stride)
  total = max((Lout - 1) * stride + K - L, 0)
  left = total // 2
  y = np.zeros((B, Lout, Cout), np.float64)
  for o in range(Lout):
    for k in range(K):
      i = stride * o + k - left
      if 0 <= i < L:
        y[:, o, :] += x[:, i, :].astype(np.float64) @ kernel[k].astype(np.float64)
  return y + bias


def naive_conv1d_transpose_same(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray, stride=2) -> np.ndarray:
  B, Lin, Cin = x.shape
  K, _, Cout, _ = kernel.shape
  Lout = Lin * stride
  left = max(K - stride, 0) // 2
  y = np.zeros((B, Lout, Cout), np.float64)
  for i in range(Lin):
    for k in range(K):
      o = stride * i + k - left
      if 0 <= o < Lout:
        y[:, o, :] += x[:, i, :].astype(np.float64) @ kernel[k, 0].astype(np.float64).T
  return y + bias


# ----------------------------------------------------------------------------- synthetic inputs

def synthetic_batch(hp: HParams, batch: int, seed: int = 1234, n_critic: Optional[int] = None):
  """Deterministic inputs shared by oracle, tests and bench (numpy RandomState so the GPU box
  and this container agree bit-for-bit)."""
  n_critic = hp.n_critic if n_critic is None else n_critic
  rng = np.random.RandomState(seed)
  L, C = hp.signal_shape
  real = rng.uniform(0.0, 1.0, size=(batch, L, C)).astype(np.float32)
  noises = rng.standard_normal((n_critic + 1, batch, hp.noise_dim)).astype(np.float32)
  alphas = rng.uniform(0.0, 1.0, size=(n_critic, batch)).astype(np.float32)
  shifts = rng.randint(-hp.m, hp.m + 1, size=(12 * n_critic + 4,)).astype(np.int32)
  return real, noises, alphas, shifts


# ----------------------------------------------------------------------------- mixed-precision restatement
# The reference's --mixed_precision runs Keras' mixed_float16 policy (main.py:22-30): variables stay fp32,
# every layer computes in and emits the 16-bit compute dtype, the final activations are fp32
# (calciumgan.py:98-101,190).  The B200 engine's policy is the bf16 analogue: fp32 master weights,
# bf16 copies of the conv / per-timestep-dense kernels, fp32 accumulation, and every *stored* activation
# or activation-gradient rounded to bf16 once (the storage points are listed at each R(...) below).
# Because LeakyReLU's slope is discontinuous at 0, any 16-bit policy flips the slope of pre-activations
# that round across 0, so its gradients differ from an fp32/fp64 evaluation by O(sqrt(ulp)), not O(ulp)
# (measured: ~6e-2 for bf16, ~2e-2 for fp16 at random init).  The functions below restate the engine's
# policy so the CUDA path can be checked at its own rounding points.

def round_bf16(x):
  return x.to(torch.float32).to(torch.bfloat16).to(x.dtype)


class _RoundST(torch.autograd.Function):
  """Round-to-bf16 on the way forward AND on the gradient coming back (storage rounding of both)."""

  @staticmethod
  def forward(ctx, x):
    return round_bf16(x)

  @staticmethod
  def backward(ctx, g):
    return _RoundST.apply(g)


def _R(x):
  return _RoundST.apply(x)


def generator_forward_mixed(gw, noise, hp: HParams, R=None):
  """generator_forward with the engine's bf16 storage points (dense0 / LN params / biases stay fp32)."""
  _R = R if R is not None else _RoundST.apply
  w, nd = calculate_noise_shape(hp.signal_shape, hp.noise_dim, NUM_LAYERS, hp.strides)
  x = _R(leaky_relu(noise @ gw[0] + gw[1])).reshape(noise.shape[0], w, nd)
  i = 2
  for l in range(NUM_LAYERS):
    x = _R(conv1d_transpose_same(x, _R(gw[i]), gw[i + 1], hp.strides))   # AG_l
    i += 2
    if hp.layer_norm:
      x = layer_norm(x, gw[i], gw[i + 1])
      i += 2
    x = _R(leaky_relu(x))                                                # HG_l
  x = x @ _R(gw[i]) + gw[i + 1]
  return torch.sigmoid(x) if hp.normalize else x                        # fp32 output


def discriminator_forward_mixed(dw, x, shifts, hp: HParams):
  x = _R(x)                                                              # X_0
  for l in range(NUM_LAYERS):
    x = _R(leaky_relu(conv1d_same(x, _R(dw[2 * l]), dw[2 * l + 1], hp.strides)))   # H_l
    if l < NUM_LAYERS - 1:
      x = phase_shuffle(x, int(shifts[l]))
  return x.reshape(x.shape[0], -1) @ dw[10] + dw[11]                     # fp32 head


def generator_step_mixed(gw, dw, real, noise, shifts, hp: HParams, dtype=torch.float64):
  """generator_step under the bf16 storage policy (autograd with straight-through rounding)."""
  gw = [_t(a, dtype).clone().requires_grad_(True) for a in gw]
  dw = [_t(a, dtype) for a in dw]
  fake = generator_forward_mixed(gw, _t(noise, dtype), hp)
  fake_out = discriminator_forward_mixed(dw, fake, shifts, hp)
  loss = -fake_out.mean()
  grads = torch.autograd.grad(loss, gw)
  return {'gen_loss': float(loss.detach()), 'fake': fake.detach(), 'fake_out': fake_out.detach(),
          'grads': [x.detach() for x in grads]}


def critic_step_mixed(gw, dw, real, noise, alpha, shifts, hp: HParams, dtype=torch.float64, R=round_bf16):
  """critic_step under the bf16 storage policy, written in the engine's own pass order
  (concatenated [real; fake; xhat] batch, 4-pass gradient penalty, one weight-gradient pass)."""
  gw = [_t(a, dtype) for a in gw]
  dwm = [_t(a, dtype) for a in dw]
  dwq = [R(w) if i < 10 and w.ndim > 1 else w for i, w in enumerate(dwm)]
  real, noise, alpha = _t(real, dtype), _t(noise, dtype), _t(alpha, dtype).reshape(-1, 1, 1)
  with torch.no_grad():
    fake = generator_forward_mixed(gw, noise, hp, R)
  B, K = real.shape[0], hp.kernel_size
  sh = np.asarray(shifts).reshape(3, 4)
  X, H = [R(torch.cat([real, fake, alpha * real + (1 - alpha) * fake]))], [None]

  def per_group(fn, x, l):
    return torch.cat([fn(x[g * B:(g + 1) * B], int(sh[g][l])) for g in range(3)])

  for l in range(NUM_LAYERS):
    h = R(leaky_relu(conv1d_same(X[l], dwq[2 * l], dwm[2 * l + 1])))
    H.append(h)
    X.append(per_group(phase_shuffle, h, l) if l < NUM_LAYERS - 1 else h)
  scores = X[5].reshape(3 * B, -1) @ dwm[10] + dwm[11]
  coef = torch.cat([torch.full((B,), -1.0 / B, dtype=dtype), torch.full((B,), 1.0 / B, dtype=dtype),
                    torch.ones(B, dtype=dtype)]).reshape(-1, 1, 1)
  slope = lambda h: torch.where(h > 0, torch.ones_like(h), torch.full_like(h, LEAKY_ALPHA))
  DA = [None] * (NUM_LAYERS + 1)
  DA[5] = R(slope(H[5]) * coef * dwm[10].reshape(1, H[5].shape[1], H[5].shape[2]))
  for l in range(NUM_LAYERS, 1, -1):
    dx = R(conv1d_same_dgrad(DA[l], dwq[2 * (l - 1)], X[l - 1].shape[1]))
    DA[l - 1] = R(slope(H[l - 1]) * per_group(_ps_transpose, dx, l - 2))
  g = R(conv1d_same_dgrad(DA[1][2 * B:], dwq[0], real.shape[1]))
  n = torch.sqrt((g.reshape(B, -1)**2).sum(1))
  gp = ((n - 1.0)**2).mean()
  V = [R(hp.gradient_penalty * (2.0 / B) * ((n - 1.0) / n).reshape(B, 1, 1) * g)]
  for l in range(NUM_LAYERS):
    c = R(slope(H[l + 1][2 * B:]) * conv1d_same(V[l], dwq[2 * l], None))
    V.append(phase_shuffle(c, int(sh[2][l])) if l < NUM_LAYERS - 1 else c)
  grads = []
  for l in range(NUM_LAYERS):
    grads.append(conv1d_same_wgrad(torch.cat([X[l][:2 * B], V[l]]), DA[l + 1], K))
    grads.append(DA[l + 1][:2 * B].sum((0, 1)))
  grads.append((coef * torch.cat([X[5][:2 * B], V[5]])).sum(0).reshape(-1, 1))
  grads.append(torch.zeros(1, dtype=dtype))
  real_loss, fake_loss = -scores[:B].mean(), scores[B:2 * B].mean()
  return {'dis_loss': float(real_loss + fake_loss + hp.gradient_penalty * gp), 'gradient_penalty': float(gp),
          'fake': fake, 'real_out': scores[:B], 'fake_out': scores[B:2 * B], 'gp_grad': g, 'gp_norm': n,
          'grads': grads, 'H': H, 'X': X}
