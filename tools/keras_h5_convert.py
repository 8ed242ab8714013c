#!/usr/bin/env python
"""Converts between the reference's Keras HDF5 weight files (ModelCheckpoint(save_weights_only=True),
src/utils/KerasCallbacks.py:54-61; model.load_weights, src/models/predict_model.py:76) and the .npz container
RvipUNet.save_weights / load_weights use.  Needs h5py -- which is NOT installed in the B200 image, so run this once on any
machine that has it (the reference's own conda environment does: environment.yml:49 h5py==2.10.0).

  python tools/keras_h5_convert.py to-npz model.h5 model.h5.npz      # reference-trained weights -> this framework
  python tools/keras_h5_convert.py to-h5  model.h5.npz model.h5      # weights trained here -> reference (tf.keras 2.3)

Both containers use the same keys: top-level `layer_names`; per layer `<layer>/<layer>/<var>:0` with var in kernel, bias,
gamma, beta, moving_mean, moving_variance.  Layers are matched by POSITION on load (Keras' auto-numbered
names depend on how many layers the writing process had created before)."""
import sys

import numpy as np


def to_npz(h5_path, npz_path):
    import h5py
    with h5py.File(h5_path, 'r') as f:
        g = f['model_weights'] if 'model_weights' in f else f
        names = [n.decode() if isinstance(n, bytes) else n for n in g.attrs['layer_names']]
        blob, kept = {}, []
        for n in names:
            wn = [w.decode() if isinstance(w, bytes) else w for w in g[n].attrs['weight_names']]
            if not wn:
                continue
            kept.append(n)
            for w in wn:
                blob[w] = np.asarray(g[n][w])
        blob['layer_names'] = np.array(kept)
    np.savez(npz_path, **blob)


def to_h5(npz_path, h5_path):
    import h5py
    z = np.load(npz_path, allow_pickle=False)
    names = [str(n) for n in z['layer_names']]
    with h5py.File(h5_path, 'w') as f:
        f.attrs['layer_names'] = [n.encode() for n in names]
        f.attrs['backend'] = b'tensorflow'
        f.attrs['keras_version'] = b'2.4.0'
        for n in names:
            grp = f.create_group(n)
            keys = [k for k in z.files if k.startswith(n + '/' + n + '/')]
            order = ['kernel:0', 'bias:0', 'gamma:0', 'beta:0', 'moving_mean:0', 'moving_variance:0']
            keys.sort(key=lambda k: order.index(k.rsplit('/', 1)[-1]))
            grp.attrs['weight_names'] = [k.encode() for k in keys]
            for k in keys:
                grp.create_dataset(k, data=z[k])


if __name__ == '__main__':
    if len(sys.argv) != 4 or sys.argv[1] not in ('to-npz', 'to-h5'):
        sys.exit(__doc__)
    (to_npz if sys.argv[1] == 'to-npz' else to_h5)(sys.argv[2], sys.argv[3])
