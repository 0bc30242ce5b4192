"""One name -> constructor table for both plugin kinds of the reference (models: gan/models/registry.py:6-13, algorithms:
gan/algorithms/registry.py:4-11). What a caller of the reference relies on is kept: `@register('name')` returns the
decorated callable unchanged, and asking for an unknown name prints the reference's message and leaves through
`exit()` (SystemExit) instead of raising KeyError."""
import sys


class PluginTable(object):

  def __init__(self, missing_message):
    self._missing_message = missing_message      # format string with one slot for the requested name
    self._entries = {}

  def register(self, name):
    """decorator: file `plugin` under `name` (a later registration of the same name replaces the earlier one)"""

    def bind(plugin):
      self._entries[name] = plugin
      return plugin

    return bind

  def names(self):
    return sorted(self._entries)

  def __contains__(self, name):
    return name in self._entries

  def resolve(self, name):
    plugin = self._entries.get(name)
    if plugin is None:
      print(self._missing_message.format(name))
      sys.exit()
    return plugin
