"""Writes tests/golden/keras_weights_tiny.h5 (+ .npz with the same arrays) with utils/hdf5_lite.py: a four-layer Keras
weight file small enough to inspect by hand.  On a machine with h5py, `python tools/keras_h5_convert.py verify
tests/golden/keras_weights_tiny.h5` opens it with the HDF5 library and compares every array with what hdf5_lite reads --
the independent check of the WRITER that this image cannot run (no HDF5 library here)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cmr_landmark_detection_b200.utils import hdf5_lite as H  # noqa: E402


def layers():
    rng = np.random.default_rng(2024)
    return [('conv2d', [('conv2d/kernel:0', rng.standard_normal((3, 3, 1, 4)).astype(np.float32)),
                        ('conv2d/bias:0', rng.standard_normal(4).astype(np.float32))]),
            ('batch_normalization', [('batch_normalization/%s:0' % n, rng.standard_normal(4).astype(np.float32))
                                     for n in ('gamma', 'beta', 'moving_mean', 'moving_variance')]),
            ('max_pooling2d', []),
            ('unet', [('unet/kernel:0', rng.standard_normal((1, 1, 4, 2)).astype(np.float32)),
                      ('unet/bias:0', np.zeros(2, np.float32))])]


if __name__ == '__main__':
    here = os.path.dirname(os.path.abspath(__file__))
    H.save_keras_weights(os.path.join(here, 'keras_weights_tiny.h5'), layers())
    np.savez(os.path.join(here, 'keras_weights_tiny.npz'), **{'%s|%s' % (l, w): a for l, ws in layers() for w, a in ws})
