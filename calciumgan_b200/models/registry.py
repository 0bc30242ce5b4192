"""Model plugins: `@register(name)` on a builder `fn(hparams) -> (generator, discriminator)` and `get_models`, the surface
of the reference's gan/models/registry.py:6-33 (unknown name: message + exit; parameter counts logged under
model/trainable_parameters/*; `.summary()` of both models when hparams.verbose)."""
from ..plugin_registry import PluginTable
from .utils import count_trainable_params

_table = PluginTable('models {} not found')
register = _table.register


def get_models(hparams, summary=None):
  generator, discriminator = _table.resolve(hparams.model)(hparams)
  if summary is not None:
    for role, model in (('generator', generator), ('discriminator', discriminator)):
      summary.scalar('model/trainable_parameters/' + role, count_trainable_params(model))
  if getattr(hparams, 'verbose', 0):
    generator.summary()
    print('')
    discriminator.summary()
  return generator, discriminator
