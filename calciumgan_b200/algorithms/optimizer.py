"""Optimizer facade with the reference's interface (gan/algorithms/optimizer.py:5-34).

Adam itself (Keras form, lr_t = lr*sqrt(1-b2^t)/(1-b1^t), eps 1e-7 outside the sqrt) is the
fused CUDA kernel behind cg_apply_update; fp32 master weights and moments live in the engine.
Under --mixed_precision the compute type is bf16, whose exponent range equals fp32, so the
reference's dynamic loss scaling (optimizer.py:10-12,23-29) is kept as an identity (scale 1).
"""
from .. import _lib as L


class Optimizer(object):

  def __init__(self, hparams, engine=None, which=L.GENERATOR):
    self._mixed_precision = getattr(hparams, 'mixed_precision', False)
    self._engine, self._which = engine, which
    self.learning_rate = hparams.learning_rate
    self.loss_scale = 1.0

  @property
  def iterations(self):
    return self._engine.get_step(self._which)

  @iterations.setter
  def iterations(self, value):
    # checkpoints written by the reference hold tf.Variable objects; accept anything int()-able
    value = value.numpy() if hasattr(value, 'numpy') else value
    self._engine.set_step(self._which, int(value))

  def get_scaled_loss(self, loss):
    return loss

  def get_unscaled_gradients(self, scaled_gradients):
    return scaled_gradients

  def update(self, model=None, scaled_loss=None, tape=None):
    """optimizer.py:31-34: gradients already sit in the engine's flat buffer."""
    self._engine.apply_update(self._which)
