"""A torch-backed stand-in for the handful of TensorFlow 2.3 / Keras calls that the reference's WGAN-GP hot path makes.

TEST INFRASTRUCTURE ONLY (see oracle/reference_runner.py). TensorFlow 2.3.1 cannot be installed in this image, so the
reference's own files (gan/models/calciumgan.py, gan/models/utils.py, gan/algorithms/{gan,wgan_gp,optimizer}.py,
gan/utils/signals_metrics.py, gan/utils/utils.py) are executed UNMODIFIED from /root/reference over this module: the
layer order, shapes, hyper-parameters, PhaseShuffle logic, loss composition, tape structure and update order are then
the reference's, and only the semantics of the TF ops listed here are restated (float64 torch; the Keras defaults they
rely on are spelled out at each definition). Everything not needed by that path raises.
"""
import numpy as np
import torch

from . import keras  # noqa: F401  (tf.keras.*)

float16, float32, float64, int32 = 'float16', 'float32', 'float64', 'int32'
DTYPE = torch.float64


def _t(x):
  if isinstance(x, torch.Tensor):
    return x
  if isinstance(x, keras.Variable):
    return x.t
  return torch.as_tensor(np.asarray(x), dtype=DTYPE)


def function(fn=None, **_kw):          # tf.function: eager here
  return fn if fn is not None else (lambda f: f)


# ---- reductions / elementwise (wgan_gp.py:20,49-50,58-59,92-93; signals_metrics.py:9-28) ----
def reduce_mean(x, axis=None):
  if isinstance(x, (list, tuple)):
    x = torch.stack([_t(v) for v in x])
  x = _t(x)
  return x.mean() if axis is None else x.mean(dim=axis)


def reduce_min(x, axis=None):
  x = _t(x)
  return x.min() if axis is None else x.min(dim=axis).values


def reduce_max(x, axis=None):
  x = _t(x)
  return x.max() if axis is None else x.max(dim=axis).values


def square(x):
  return _t(x) ** 2


def norm(x, axis=None):                # tf.norm: sqrt(sum x^2), no epsilon
  x = _t(x)
  return torch.sqrt((x * x).sum()) if axis is None else torch.sqrt((x * x).sum(dim=axis))


def reshape(x, shape):
  return _t(x).reshape(tuple(int(s) for s in shape))


def expand_dims(x, axis):
  return _t(x).unsqueeze(axis)


def squeeze(x, axis=None):
  return _t(x).squeeze() if axis is None else _t(x).squeeze(axis)


def ones_like(x):
  return torch.ones_like(_t(x))


def zeros_like(x):
  return torch.zeros_like(_t(x))


def ensure_shape(x, shape):            # the batch dimension of a Keras symbolic shape is None
  got, want = tuple(x.shape)[1:], tuple(shape)[1:]
  if got != tuple(int(s) for s in want):
    raise ValueError('ensure_shape: %s vs %s' % (got, want))
  return x


def pad(x, paddings, mode='CONSTANT'):
  """tf.pad on a rank-3 tensor; REFLECT excludes the border like numpy's 'reflect' (calciumgan.py:135)."""
  x = _t(x)
  (b0, b1), (t0, t1), (c0, c1) = [(int(a), int(b)) for a, b in paddings]
  assert (b0, b1, c0, c1) == (0, 0, 0, 0), 'only the time axis is padded on this path'
  if mode.upper() == 'REFLECT':
    w = x.shape[1]
    assert t0 < w and t1 < w
    idx = list(range(t0, 0, -1)) + list(range(w)) + list(range(w - 2, w - 2 - t1, -1))
    return x[:, idx, :]
  assert mode.upper() == 'CONSTANT'
  return torch.nn.functional.pad(x, (0, 0, t0, t1))


class _Math(object):
  @staticmethod
  def abs(x):
    return abs(x) if isinstance(x, (int, np.integer)) else torch.abs(_t(x))

  @staticmethod
  def reduce_std(x, axis=None):        # population standard deviation
    x = _t(x)
    return x.std(unbiased=False) if axis is None else x.std(dim=axis, unbiased=False)


math = _Math()


# ---- randomness: draws are injected so that the reference code and the oracle see identical numbers ----
class _Random(object):
  def __init__(self):
    self.normal_q, self.uniform_q, self.int_q = [], [], []
    self.log = []

  def inject(self, normal=(), uniform=(), ints=()):
    self.normal_q, self.uniform_q, self.int_q = list(normal), list(uniform), list(ints)
    self.log = []

  def normal(self, shape, **_kw):
    v = _t(self.normal_q.pop(0))
    assert tuple(v.shape) == tuple(shape), (v.shape, shape)
    self.log.append('normal%s' % (tuple(shape),))
    return v

  def uniform(self, shape, minval=0, maxval=None, dtype=float32, **_kw):
    if keras._STATE['chain'] is not None:   # functional-API construction traces call() on the probe tensor: no draw
      return 0 if dtype == int32 else torch.zeros(tuple(shape), dtype=DTYPE)
    if dtype == int32:
      assert list(shape) == []
      v = int(self.int_q.pop(0))
      assert minval <= v < maxval, (minval, v, maxval)
      self.log.append('int')
      return v
    v = _t(self.uniform_q.pop(0)).reshape(tuple(shape))
    self.log.append('uniform%s' % (tuple(shape),))
    return v

  def set_seed(self, _seed):
    pass


random = _Random()


# ---- tf.GradientTape over torch autograd: gradient of a non-scalar target is the gradient of its sum; the outer tape
# of _train_discriminator differentiates through the inner tape's gradient (create_graph) ----
class GradientTape(object):
  def __init__(self, persistent=False):
    pass

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    return False

  def watch(self, x):
    if not x.requires_grad:
      x.requires_grad_(True)

  def gradient(self, target, sources):
    single = not isinstance(sources, (list, tuple))
    srcs = [sources] if single else list(sources)
    ts = [s.t if isinstance(s, keras.Variable) else s for s in srcs]
    target = _t(target)
    g = torch.autograd.grad(target.sum(), ts, create_graph=True, retain_graph=True, allow_unused=True)
    g = [torch.zeros_like(t) if gi is None else gi for gi, t in zip(g, ts)]
    return g[0] if single else g


def py_function(func, inp, Tout):
  raise NotImplementedError('tf.py_function is outside the WGAN-GP hot path')
