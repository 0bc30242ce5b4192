"""Helpers with the names of the reference's gan/models/utils.py (only what the hot path uses)."""
import numpy as np


def count_trainable_params(model):
  ''' return the number of trainable parameters (gan/models/utils.py:11-14) '''
  return int(np.sum([int(np.prod(p.shape)) for p in model.trainable_weights]))
