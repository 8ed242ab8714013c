"""CPU: the connected-component oracle against golden vectors produced by the reference's own filter (cv2)."""
import os

import numpy as np

from oracle import cc_ref

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'cc_golden.npz')


def test_cc_oracle_matches_reference_golden():
    g = np.load(GOLD)
    for c in [str(c) for c in g['cases']]:
        vol = g[c + '_in']
        if vol.shape[1] > 64:
            vol = vol[:2]          # the pure-Python union-find is slow; two slices of the large case suffice here
        out, ties = cc_ref.clean_2d_cc(vol, return_ties=True)
        ref = g[c + '_out'][:len(vol)]
        for z in range(len(vol)):
            if not ties[z]:
                assert np.array_equal(out[z], ref[z]), (c, z)
            else:                 # a tied maximum is decided by OpenCV's internal label numbering (see cc_ref):
                for val in (1, 2):   # the kept component must at least have the same (maximal) area
                    assert (out[z] == val).sum() == (ref[z] == val).sum(), (c, z, val)
