"""Convert the HDF5-free store of calciumgan_b200/utils/h5_helper.py (<file>.h5.parts/) to the reference's HDF5 layout.
Needs h5py (not available in the build image):  python tools/parts_to_h5.py runs/generated/epoch019_signals.h5"""
import os
import re
import sys
from glob import glob

import numpy as np


def main(filename):
  import h5py
  parts = filename + '.parts'
  names = sorted({re.match(r'(.+)\.\d+\.npy$', os.path.basename(p)).group(1) for p in glob(os.path.join(parts, '*.npy'))})
  with h5py.File(filename, mode='w') as file:
    for name in names:
      blocks = sorted(glob(os.path.join(parts, name + '.*.npy')))
      first = np.load(blocks[0])
      ds = file.create_dataset(name, shape=first.shape, dtype=first.dtype, data=first, chunks=True, maxshape=(None,) + first.shape[1:])
      for path in blocks[1:]:
        v = np.load(path)
        ds.resize((ds.shape[0] + v.shape[0]), axis=0)
        ds[-v.shape[0]:] = v
      print(name, ds.shape, ds.dtype)


if __name__ == '__main__':
  main(sys.argv[1])
