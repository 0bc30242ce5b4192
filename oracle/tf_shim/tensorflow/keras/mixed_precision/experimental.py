class LossScaleOptimizer(object):      # optimizer.py:10-12, only with --mixed_precision
  def __init__(self, *a, **k):
    raise NotImplementedError('loss scaling is not modelled: the shim runs the fp32 policy in float64')
