"""Keras layers used by gan/models/calciumgan.py and gan/models/utils.py, with the TF 2.3 defaults they rely on."""
import numpy as np
import torch
import torch.nn.functional as F

from . import DTYPE, Variable, _STATE

_COUNTS = {}


def _uid(prefix):
  n = _COUNTS.get(prefix, 0)
  _COUNTS[prefix] = n + 1
  return prefix if n == 0 else '%s_%d' % (prefix, n)


def _glorot(shape, fan_in, fan_out):
  lim = np.sqrt(6.0 / (fan_in + fan_out))
  return np.random.uniform(-lim, lim, size=shape)


class Layer(object):
  def __init__(self, **_kw):
    self.name = _uid(type(self).__name__.lower())
    self.built = False
    self._own = []

  def build(self, input_shape):
    pass

  def call(self, inputs):
    raise NotImplementedError

  def add(self, value, suffix):
    v = Variable(value, '%s/%s:0' % (self.name, suffix))
    self._own.append(v)
    return v

  def __call__(self, inputs, **_kw):
    if _STATE['chain'] is not None and _STATE['depth'] == 0:
      _STATE['chain'].append(self)
    if not self.built:
      self.build(tuple(inputs.shape))
      self.built = True
    _STATE['depth'] += 1
    try:
      return self.call(inputs)
    finally:
      _STATE['depth'] -= 1

  @property
  def trainable_variables(self):
    out = list(self._own)
    for value in self.__dict__.values():   # sub-layers in attribute-assignment order (Keras tracking order)
      if isinstance(value, Layer):
        out += value.trainable_variables
    return out


class Dense(Layer):                        # kernel (in, units) glorot-uniform, bias zeros; acts on the last axis
  def __init__(self, units, **kw):
    super().__init__()
    self.units = int(units)

  def build(self, shape):
    self.kernel = self.add(_glorot((shape[-1], self.units), shape[-1], self.units), 'kernel')
    self.bias = self.add(np.zeros(self.units), 'bias')

  def call(self, x):
    return x @ self.kernel.t + self.bias.t


class LeakyReLU(Layer):                    # Keras default alpha = 0.3
  def __init__(self, alpha=0.3, **kw):
    super().__init__()
    self.alpha = alpha

  def call(self, x):
    return torch.where(x > 0, x, self.alpha * x)


class Activation(Layer):
  def __init__(self, activation, dtype=None, **kw):
    super().__init__()
    self.activation = activation

  def call(self, x):
    if self.activation == 'linear':
      return x
    if self.activation == 'sigmoid':
      return torch.sigmoid(x)
    raise NotImplementedError(self.activation)


class Reshape(Layer):
  def __init__(self, target_shape, **kw):
    super().__init__()
    self.target_shape = tuple(int(s) for s in target_shape)

  def call(self, x):
    return x.reshape((x.shape[0],) + self.target_shape)


class Flatten(Layer):                      # row-major over (time, channels)
  def call(self, x):
    return x.reshape(x.shape[0], -1)


class LayerNormalization(Layer):           # axis -1, epsilon 1e-3, biased variance, gamma 1 / beta 0
  def __init__(self, axis=-1, epsilon=1e-3, **kw):
    super().__init__()
    assert axis == -1
    self.epsilon = epsilon

  def build(self, shape):
    self.gamma = self.add(np.ones(shape[-1]), 'gamma')
    self.beta = self.add(np.zeros(shape[-1]), 'beta')

  def call(self, x):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + self.epsilon) * self.gamma.t + self.beta.t


class BatchNormalization(Layer):
  def __init__(self, **kw):
    raise NotImplementedError('--batch_norm is outside the supported path')


class Conv1D(Layer):
  """kernel (K, Cin, Cout); padding='same': Lout = ceil(L / s), pad_total = max((Lout - 1) s + K - L, 0),
  left = pad_total // 2 (the extra element goes to the right); cross-correlation."""

  def __init__(self, filters, kernel_size, strides=1, padding='valid', **kw):
    super().__init__()
    assert padding == 'same'
    self.filters, self.k, self.s = int(filters), int(kernel_size), int(strides)

  def build(self, shape):
    cin = shape[-1]
    self.kernel = self.add(_glorot((self.k, cin, self.filters), self.k * cin, self.k * self.filters), 'kernel')
    self.bias = self.add(np.zeros(self.filters), 'bias')

  def call(self, x):
    L = x.shape[1]
    lout = -(-L // self.s)
    total = max((lout - 1) * self.s + self.k - L, 0)
    left = total // 2
    xp = F.pad(x.permute(0, 2, 1), (left, total - left))
    y = F.conv1d(xp, self.kernel.t.permute(2, 1, 0), self.bias.t, stride=self.s)
    return y.permute(0, 2, 1)


class Conv2DTranspose(Layer):
  """Only the (K, 1) / (s, 1) form of models/utils.py:79-84. kernel (K, 1, Cout, Cin); padding='same': output length
  L s; the full transposed convolution (length (L - 1) s + K) is cropped by pad_total = K - s with
  left = pad_total // 2 (conv_utils.deconv_output_length + the gradient-of-SAME-conv definition)."""

  def __init__(self, filters, kernel_size, strides, padding='valid', output_padding=None, **kw):
    super().__init__()
    assert padding == 'same' and output_padding is None
    assert kernel_size[1] == 1 and strides[1] == 1
    self.filters, self.k, self.s = int(filters), int(kernel_size[0]), int(strides[0])

  def build(self, shape):
    cin = shape[-1]
    self.kernel = self.add(_glorot((self.k, 1, self.filters, cin), self.k * cin, self.k * self.filters), 'kernel')
    self.bias = self.add(np.zeros(self.filters), 'bias')

  def call(self, x):                       # x (B, L, 1, Cin)
    B, L = x.shape[0], x.shape[1]
    w = self.kernel.t[:, 0].permute(2, 1, 0)            # (Cin, Cout, K)
    full = F.conv_transpose1d(x[:, :, 0, :].permute(0, 2, 1), w, stride=self.s)   # length (L - 1) s + K
    total = max(self.k - self.s, 0)
    left = total // 2
    y = full[:, :, left:left + L * self.s] + self.bias.t[None, :, None]
    return y.permute(0, 2, 1).unsqueeze(2)
