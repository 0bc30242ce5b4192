"""Dataset store behind `save_fake_signals` / `cache_validation_set`, with the call surface of the reference's
gan/utils/h5_helper.py: `write` (create-or-append along axis 0), `get` (whole dataset, one neuron, or one trial; NWC
arrays), `overwrite`, `get_dataset_length`, `contains`.

Two backends implement it. `H5pyStore` is used whenever h5py can be imported: HDF5 files with the reference's layout
(h5_helper.py:13-30: one chunked dataset per key, resizable along axis 0), so the reference's analysis scripts read them
unchanged. This build image has no HDF5 library at all (no h5py, PyTables, netCDF4 or libhdf5) and a hand-assembled HDF5
file could not be checked against libhdf5 here, so the HDF5-free fallback does not pretend to be one:
`PartsStore` keeps `<filename>.parts/<key>.<index>.npy`, one file per appended block -- appends cost O(block), reads
concatenate. `tools/parts_to_h5.py` converts such a directory to the reference's HDF5 layout on a machine with h5py."""
import os
import re
from glob import glob

import numpy as np

try:
  import h5py
except ImportError:
  h5py = None


def _missing(name):
  return KeyError('{} cannot be found'.format(name))


def parts_dir(filename):
  return filename + '.parts'


class H5pyStore(object):
  name = 'h5py'

  @staticmethod
  def exists(filename):
    return os.path.exists(filename)

  @staticmethod
  def _dataset(handle, name):
    if name not in handle:
      raise _missing(name)
    return handle[name]

  def append(self, filename, name, block):
    with h5py.File(filename, 'a') as handle:
      if name not in handle:
        handle.create_dataset(name, data=block, chunks=True, maxshape=(None,) + block.shape[1:])
        return
      ds = handle[name]
      old = ds.shape[0]
      ds.resize(old + block.shape[0], axis=0)
      ds[old:] = block

  def replace(self, filename, name, value):
    with h5py.File(filename, 'r+') as handle:
      self._dataset(handle, name)
      del handle[name]
      handle.create_dataset(name, data=value)

  def read(self, filename, name, neuron, trial):
    with h5py.File(filename, 'r') as handle:
      ds = self._dataset(handle, name)
      if neuron is not None:
        return ds[:, :, neuron]
      return ds[trial] if trial is not None else ds[()]

  def length(self, filename, name):
    with h5py.File(filename, 'r') as handle:
      return int(self._dataset(handle, name).shape[0])

  def has(self, filename, name):
    with h5py.File(filename, 'r') as handle:
      return name in handle


class PartsStore(object):
  name = 'npy-parts'

  @staticmethod
  def exists(filename):
    return os.path.isdir(parts_dir(filename))

  @staticmethod
  def _blocks(filename, name, required=True):
    """paths of the blocks of dataset `name`, in append order"""
    pattern = re.compile(re.escape(name) + r'\.(\d+)\.npy$')
    numbered = []
    for path in glob(os.path.join(parts_dir(filename), name + '.*.npy')):
      m = pattern.match(os.path.basename(path))
      if m:
        numbered.append((int(m.group(1)), path))
    if required and not numbered:
      raise _missing(name)
    return [path for _, path in sorted(numbered)]

  @staticmethod
  def _block_path(filename, name, index):
    return os.path.join(parts_dir(filename), '%s.%06d.npy' % (name, index))

  def append(self, filename, name, block):
    os.makedirs(parts_dir(filename), exist_ok=True)
    block = np.asarray(block)
    blocks = self._blocks(filename, name, required=False)
    if blocks:
      first = np.load(blocks[0], mmap_mode='r')
      if first.shape[1:] != block.shape[1:] or first.dtype != block.dtype:
        raise ValueError('cannot append %s %s to dataset %s of %s %s' %
                         (block.dtype, block.shape[1:], name, first.dtype, first.shape[1:]))
    np.save(self._block_path(filename, name, len(blocks)), block)

  def replace(self, filename, name, value):
    for path in self._blocks(filename, name):
      os.remove(path)
    np.save(self._block_path(filename, name, 0), np.asarray(value))

  def read(self, filename, name, neuron, trial):
    blocks = self._blocks(filename, name)
    if trial is not None:      # only the block that holds the trial is read
      if trial < 0:
        trial += self.length(filename, name)
      for path in blocks:
        block = np.load(path, mmap_mode='r')
        if trial < len(block):
          return np.array(block[trial])
        trial -= len(block)
      raise IndexError('trial out of range')
    if neuron is not None:
      return np.concatenate([np.load(path, mmap_mode='r')[:, :, neuron] for path in blocks], axis=0)
    return np.concatenate([np.load(path) for path in blocks], axis=0)

  def length(self, filename, name):
    return int(sum(len(np.load(path, mmap_mode='r')) for path in self._blocks(filename, name)))

  def has(self, filename, name):
    return bool(self._blocks(filename, name, required=False))


_store = H5pyStore() if h5py is not None else PartsStore()


def backend():
  return _store.name


def exists(filename):
  """the store `filename` has been written to"""
  return _store.exists(filename)


def write(filename, content):
  """create each dataset of `content` (dict name -> NWC array) or append to it along axis 0 (h5_helper.py:13-30)"""
  assert type(content) == dict
  for name, block in content.items():
    _store.append(filename, name, block)


def overwrite(filename, name, value):
  """replace dataset `name`; KeyError when it does not exist (h5_helper.py:33-39)"""
  _store.replace(filename, name, value)


def get(filename, name, neuron=None, trial=None):
  """dataset `name`, or one neuron of it ([:, :, neuron]), or one trial ([trial]) (h5_helper.py:42-60)"""
  assert neuron is None or trial is None
  return _store.read(filename, name, neuron, trial)


def get_dataset_length(filename, name):
  return _store.length(filename, name)


def contains(filename, name):
  return _store.has(filename, name)
