#!/usr/bin/env python
"""Keras HDF5 weight files (ModelCheckpoint(save_weights_only=True), src/utils/KerasCallbacks.py:54-61; model.load_weights,
src/models/predict_model.py:76)  <->  the .npz container RvipUNet.save_weights writes for paths without an .h5 suffix.

RvipUNet reads and writes `model.h5` itself (cmr_landmark_detection_b200/utils/hdf5_lite.py, no h5py needed); this tool
is for moving between the two containers and -- on a machine that HAS h5py, e.g. the reference's conda environment
(environment.yml:49 h5py==2.10.0) -- for cross-checking hdf5_lite against the HDF5 library:

  python tools/keras_h5_convert.py to-npz model.h5 weights.npz     # Keras / RvipUNet HDF5 -> .npz
  python tools/keras_h5_convert.py to-h5  weights.npz model.h5     # .npz -> Keras HDF5 (written by hdf5_lite)
  python tools/keras_h5_convert.py verify model.h5                 # needs h5py: every array and attribute read with the
                                                                   # library must equal what hdf5_lite reads

Keys of the .npz: `layer_names` (layers that carry weights, in order) and `<layer>/<weight name>` with Keras' weight names
(`conv2d/kernel:0`, `batch_normalization/gamma:0`, ...), i.e. `<layer>/<layer>/<var>:0`.  Layers are
matched by POSITION on load (Keras' auto-numbered names depend on how many layers the writing process created before)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmr_landmark_detection_b200.utils import hdf5_lite  # noqa: E402


def to_npz(h5_path, npz_path):
    layers = hdf5_lite.load_keras_weights(h5_path)
    blob = {'%s/%s' % (lname, wn): arr for lname, weights in layers for wn, arr in weights}
    blob['layer_names'] = np.array([lname for lname, _ in layers])
    np.savez(npz_path, **blob)


def to_h5(npz_path, h5_path):
    z = np.load(npz_path, allow_pickle=False)
    order = ['kernel:0', 'bias:0', 'gamma:0', 'beta:0', 'moving_mean:0', 'moving_variance:0']
    layers = []
    for lname in [str(n) for n in z['layer_names']]:
        prefix = lname + '/'
        keys = [k for k in z.files if k.startswith(prefix)]
        keys.sort(key=lambda k: order.index(k.rsplit('/', 1)[-1]) if k.rsplit('/', 1)[-1] in order else len(order))
        layers.append((lname, [(k[len(prefix):], z[k]) for k in keys]))
    hdf5_lite.save_keras_weights(h5_path, layers)


def verify(h5_path):
    import h5py
    mine = hdf5_lite.load_keras_weights(h5_path)
    n = 0
    with h5py.File(h5_path, 'r') as f:
        g = f['model_weights'] if 'model_weights' in f else f
        dec = lambda v: v.decode() if isinstance(v, bytes) else str(v)
        names = [dec(x) for x in g.attrs['layer_names']]
        kept = [x for x in names if len(g[x].attrs['weight_names'])]
        assert kept == [lname for lname, _ in mine], (kept, [lname for lname, _ in mine])
        for lname, weights in mine:
            wn = [dec(w) for w in g[lname].attrs['weight_names']]
            assert wn == [w for w, _ in weights], (lname, wn)
            for w, arr in weights:
                ref = np.asarray(g[lname][w])
                assert ref.shape == arr.shape and np.array_equal(ref, arr), (lname, w)
                n += 1
    print('%s: %d layers, %d arrays -- hdf5_lite and h5py agree' % (h5_path, len(mine), n))


if __name__ == '__main__':
    if len(sys.argv) < 3 or sys.argv[1] not in ('to-npz', 'to-h5', 'verify'):
        sys.exit(__doc__)
    if sys.argv[1] == 'verify':
        verify(sys.argv[2])
    else:
        (to_npz if sys.argv[1] == 'to-npz' else to_h5)(sys.argv[2], sys.argv[3])
