"""Generates tests/golden/cc_golden.npz by running the REFERENCE's own largest-connected-component filter
(src/data/Postprocess.py:108-120 clean_3d_prediction_2d_cc; its cv2 call passes 4 positionally, which OpenCV takes as the
`labels` slot, so it runs 8-connected -- see oracle/cc_ref.py) on
seeded label volumes.  Run in the build container only:   python tests/golden/make_cc_golden.py"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def import_postprocess():
    for m in ['SimpleITK', 'skimage', 'skimage.measure', 'skimage.exposure', 'skimage.transform', 'matplotlib',
              'matplotlib.pyplot', 'tensorflow', 'albumentations']:
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = MagicMock()
    sys.path.insert(0, '/root/reference')
    import src.data.Postprocess as P
    return P


def blobs(rng, Z, H, W, kind):
    vol = np.zeros((Z, H, W), np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    for z in range(Z):
        for lab in (1, 2):
            n = {'single': 1, 'multi': int(rng.integers(2, 5)), 'tie': 2, 'noise': 3, 'empty': 0}[kind]
            if kind == 'empty' and lab == 2:
                n = 1
            for k in range(n):
                cy, cx = rng.integers(3, H - 3), rng.integers(3, W - 3)
                r = 2 if kind == 'tie' else int(rng.integers(1, 5))
                if kind == 'tie':
                    m = (np.abs(yy - cy) <= r) & (np.abs(xx - cx) <= r)      # equal-size squares
                else:
                    m = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
                vol[z][m & (vol[z] == 0)] = lab
        if kind == 'noise':
            sp = rng.random((H, W)) < 0.04
            vol[z][sp & (vol[z] == 0)] = rng.integers(1, 3, size=int((sp & (vol[z] == 0)).sum()))
    if kind == 'multi':
        # diagonal-only contact: connected with the reference's call (8-connectivity), not under true 4-connectivity
        vol[0, 1:3, 1:3] = 0
        vol[0, 1, 1] = 1
        vol[0, 2, 2] = 1
    return vol


def main():
    P = import_postprocess()
    rng = np.random.default_rng(20211010)
    out, names = {}, []
    for i, (kind, Z, H, W) in enumerate([('single', 4, 32, 32), ('multi', 6, 48, 40), ('tie', 5, 40, 40),
                                        ('noise', 4, 64, 64), ('empty', 3, 24, 24), ('multi', 8, 128, 128)]):
        vol = blobs(rng, Z, H, W, kind)
        out['c%d_in' % i] = vol
        out['c%d_out' % i] = P.clean_3d_prediction_2d_cc(vol).astype(np.uint8)
        names.append('c%d' % i)
    out['cases'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'cc_golden.npz'), **out)
    print({n: (int((out[n + '_in'] > 0).sum()), int((out[n + '_out'] > 0).sum())) for n in names})


if __name__ == '__main__':
    main()
