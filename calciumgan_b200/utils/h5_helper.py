"""Dataset store behind `save_fake_signals` with the call surface of the reference's gan/utils/h5_helper.py
(`write` = create-or-append along axis 0, `get` with the neuron / trial selectors, `overwrite`, `get_dataset_length`,
`contains`; NWC arrays).

With h5py importable the files are HDF5 with the reference's layout (h5_helper.py:13-30: one chunked dataset per key,
resizable along axis 0), so the reference's analysis scripts read them unchanged. This image has no HDF5 library at all
(no h5py, PyTables, netCDF4 or libhdf5), and a hand-assembled HDF5 file could not be checked against libhdf5 here, so the
HDF5-free fallback does not pretend to be one: `<filename>.parts/` holds one `.npy` per appended block
(`<key>.<index>.npy`), appends cost O(block), `get` concatenates. `tools/parts_to_h5.py` converts such a directory to the
reference's HDF5 layout on a machine that has h5py."""
import os
import re
from glob import glob

import numpy as np

try:
  import h5py
except ImportError:
  h5py = None


def backend():
  return 'h5py' if h5py is not None else 'npy-parts'


def parts_dir(filename):
  return filename + '.parts'


def exists(filename):
  """the store `filename` has been written to (whichever backend holds it)"""
  return os.path.exists(filename) or os.path.isdir(parts_dir(filename))


def _blocks(filename, name):
  found = []
  for path in glob(os.path.join(parts_dir(filename), name + '.*.npy')):
    m = re.match(re.escape(name) + r'\.(\d+)\.npy$', os.path.basename(path))
    if m:
      found.append((int(m.group(1)), path))
  return [p for _, p in sorted(found)]


def write(filename, content):
  """write or append content (dict name -> array, NWC) to the store (h5_helper.py:13-30)"""
  assert type(content) == dict
  if h5py is not None:
    with h5py.File(filename, mode='a') as file:
      for k, v in content.items():
        if k in file:
          ds = file[k]
          ds.resize((ds.shape[0] + v.shape[0]), axis=0)
          ds[-v.shape[0]:] = v
        else:
          file.create_dataset(k, shape=v.shape, dtype=v.dtype, data=v, chunks=True, maxshape=(None,) + v.shape[1:])
    return
  os.makedirs(parts_dir(filename), exist_ok=True)
  for k, v in content.items():
    v = np.asarray(v)
    blocks = _blocks(filename, k)
    if blocks:
      first = np.load(blocks[0], mmap_mode='r')
      if first.shape[1:] != v.shape[1:] or first.dtype != v.dtype:
        raise ValueError('cannot append %s %s to dataset %s of %s %s' % (v.dtype, v.shape[1:], k, first.dtype, first.shape[1:]))
    np.save(os.path.join(parts_dir(filename), '%s.%06d.npy' % (k, len(blocks))), v)


def overwrite(filename, name, value):
  """replace dataset `name` (h5_helper.py:33-39); KeyError when it does not exist"""
  if h5py is not None:
    with h5py.File(filename, mode='r+') as file:
      if name not in file.keys():
        raise KeyError('{} cannot be found'.format(name))
      del file[name]
      file.create_dataset(name, shape=value.shape, dtype=value.dtype, data=value)
    return
  blocks = _blocks(filename, name)
  if not blocks:
    raise KeyError('{} cannot be found'.format(name))
  for path in blocks:
    os.remove(path)
  np.save(os.path.join(parts_dir(filename), '%s.%06d.npy' % (name, 0)), np.asarray(value))


def get(filename, name, neuron=None, trial=None):
  """the dataset `name`, or one neuron ([:, :, neuron]) / one trial ([trial]) of it (h5_helper.py:42-60)"""
  assert not (neuron is not None and trial is not None)
  if h5py is not None:
    with h5py.File(filename, mode='r') as file:
      if name not in file.keys():
        raise KeyError('{} cannot be found'.format(name))
      ds = file[name]
      if neuron is not None:
        return ds[:, :, neuron]
      if trial is not None:
        return ds[trial, :, :]
      return ds[:]
  blocks = _blocks(filename, name)
  if not blocks:
    raise KeyError('{} cannot be found'.format(name))
  if trial is not None:      # only the block that holds the trial is read
    if trial < 0:
      trial += get_dataset_length(filename, name)
    for path in blocks:
      block = np.load(path, mmap_mode='r')
      if trial < len(block):
        return np.array(block[trial])
      trial -= len(block)
    raise IndexError('trial out of range')
  if neuron is not None:
    return np.concatenate([np.load(p, mmap_mode='r')[:, :, neuron] for p in blocks], axis=0)
  return np.concatenate([np.load(p) for p in blocks], axis=0)


def get_dataset_length(filename, name):
  if h5py is not None:
    with h5py.File(filename, mode='r') as file:
      return file[name].len()
  blocks = _blocks(filename, name)
  if not blocks:
    raise KeyError('{} cannot be found'.format(name))
  return int(sum(len(np.load(p, mmap_mode='r')) for p in blocks))


def contains(filename, name):
  if h5py is not None:
    with h5py.File(filename, mode='r') as file:
      return name in list(file.keys())
  return bool(_blocks(filename, name))
