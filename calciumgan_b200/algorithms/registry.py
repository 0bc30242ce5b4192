"""Algorithm plugins: `@register(name)` on a class constructed as `cls(hparams, generator, discriminator, summary)` and
`get_algorithm`, the surface of the reference's gan/algorithms/registry.py:4-19 (unknown name: message + exit)."""
from ..plugin_registry import PluginTable

_table = PluginTable('Algorithm {} not found')
register = _table.register


def get_algorithm(hparams, generator, discriminator, summary=None):
  return _table.resolve(hparams.algorithm)(hparams, generator, discriminator, summary)
