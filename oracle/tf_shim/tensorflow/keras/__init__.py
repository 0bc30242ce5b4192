"""tf.keras subset (see ../__init__.py): functional-API models that are plain chains, Adam, the layer classes."""
import numpy as np
import torch

DTYPE = torch.float64
_STATE = {'depth': 0, 'chain': None}


class Variable(object):
  def __init__(self, value, name):
    self.t = torch.as_tensor(np.asarray(value), dtype=DTYPE).clone().requires_grad_(True)
    self.name = name

  @property
  def shape(self):
    return tuple(self.t.shape)

  def numpy(self):
    return self.t.detach().numpy().copy()

  def assign(self, value):
    with torch.no_grad():
      self.t.copy_(torch.as_tensor(np.asarray(value), dtype=DTYPE))


from . import layers  # noqa: E402
from . import mixed_precision  # noqa: E402,F401


def Input(shape, name=None):
  """Symbolic input: a batch-1 probe tensor; the layers called on it (at nesting depth 0) form the model's chain."""
  _STATE['chain'] = []
  return torch.zeros((1,) + tuple(int(s) for s in shape), dtype=DTYPE)


class Model(object):
  def __init__(self, inputs, outputs, name=None):
    self.name = name
    self.layers = list(_STATE['chain'])
    _STATE['chain'] = None
    self.output_shape = (None,) + tuple(outputs.shape)[1:]

  def __call__(self, x, training=None):
    x = torch.as_tensor(np.asarray(x), dtype=DTYPE) if not isinstance(x, torch.Tensor) else x
    for layer in self.layers:
      x = layer(x)
    return x

  @property
  def trainable_variables(self):
    out = []
    for layer in self.layers:
      out += layer.trainable_variables
    return out

  trainable_weights = trainable_variables

  def get_weights(self):
    return [v.numpy() for v in self.trainable_variables]

  def set_weights(self, weights):
    vs = self.trainable_variables
    assert len(vs) == len(weights), (len(vs), len(weights))
    for v, w in zip(vs, weights):
      assert v.shape == tuple(np.shape(w)), (v.name, v.shape, np.shape(w))
      v.assign(w)

  def summary(self):
    for v in self.trainable_variables:
      print(v.name, v.shape)


class _Backend(object):
  @staticmethod
  def count_params(p):
    return int(np.prod(p.shape))


backend = _Backend()


class _Adam(object):
  """Keras Adam (optimizer_v2/adam.py, TF 2.3): lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);
  m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2; var -= lr_t * m / (sqrt(v) + epsilon), epsilon = 1e-7."""

  def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
    self.lr, self.b1, self.b2, self.eps = float(learning_rate), beta_1, beta_2, epsilon
    self.iterations = 0
    self.slots = {}
    self.gradient_log = []     # test hook: the gradients of every apply_gradients call, in variable order

  def apply_gradients(self, grads_and_vars):
    grads_and_vars = list(grads_and_vars)
    self.gradient_log.append([g.detach().clone() for g, _ in grads_and_vars])
    self.iterations += 1
    t = self.iterations
    lr_t = self.lr * np.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
    with torch.no_grad():
      for g, var in grads_and_vars:
        g = g.detach()
        m, v = self.slots.setdefault(id(var), [torch.zeros_like(var.t), torch.zeros_like(var.t)])
        m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
        v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
        var.t.sub_(lr_t * m / (torch.sqrt(v) + self.eps))


class _Optimizers(object):
  Adam = _Adam


optimizers = _Optimizers()


class _Losses(object):
  class BinaryCrossentropy(object):    # constructed by GAN.__init__ (gan.py:27), never called by WGAN_GP
    def __init__(self, from_logits=False):
      pass

    def __call__(self, *a, **k):
      raise NotImplementedError('vanilla GAN loss is outside the WGAN-GP hot path')

  @staticmethod
  def KLD(y_true, y_pred):
    raise NotImplementedError


losses = _Losses()
