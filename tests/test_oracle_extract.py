"""Pins oracle/extract_ref.py against golden vectors produced by the REFERENCE's own functions
(tests/golden/make_extract_golden.py imports /root/reference/src/models/evaluate_cv.py)."""
import os

import numpy as np
import pytest

from oracle import extract_ref as ex


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'extract_golden.npz'))


def dec(arr):
    return [None if np.isnan(p[0]) else [float(p[0]), float(p[1])] for p in arr]


def same_pts(a, b):
    assert len(a) == len(b)
    for p, q in zip(a, b):
        if p is None or q is None:
            assert p is None and q is None
        else:
            assert p[0] == q[0] and p[1] == q[1]      # same float64 arithmetic -> exact


def test_hand_example(gold):
    a, b = ex.mean_rvip_2d(gold['hand/mask'])
    assert a == [6.0, 11.0] and b == [20.5, 6.0]
    assert list(gold['hand/ant']) == a and list(gold['hand/inf']) == b
    assert abs(ex.get_angle2x(a, b) - float(gold['hand/angle'])) < 1e-12
    assert abs(float(gold['hand/angle']) - 109.02560603756869) < 1e-9
    assert abs(ex.get_dist(a, b) - float(gold['hand/dist'])) < 1e-12


def test_label_map_and_centroids_match_reference(gold):
    for key in gold['names']:
        heat = gold[key + '/heat']
        lab = ex.label_map(heat, 0.5)
        assert np.array_equal(lab.astype(np.uint8), gold[key + '/labels']), key
        for both in (True, False):
            a, b = ex.ip_from_rvip_mask_3d(lab.astype(np.uint8), keepdim=True, both_only=both)
            same_pts(a, dec(gold[key + '/ant_both%d' % both]))
            same_pts(b, dec(gold[key + '/inf_both%d' % both]))
        a, b = ex.ip_from_rvip_mask_3d(lab.astype(np.uint8), keepdim=False, both_only=True)
        same_pts(a, dec(gold[key + '/ant_nokeep']))
        same_pts(b, dec(gold[key + '/inf_nokeep']))


def test_integer_stats_reproduce_reference_points(gold):
    """The (count, sum_row, sum_col) form the device kernel returns carries the same information."""
    for key in gold['names']:
        heat = gold[key + '/heat']
        count, srow, scol, amax, vmax = ex.extract_stats(heat, 0.5)
        for both in (True, False):
            a, b = ex.landmarks_from_stats(count, srow, scol, both_only=both, keepdim=True)
            ga, gb = dec(gold[key + '/ant_both%d' % both]), dec(gold[key + '/inf_both%d' % both])
            for p, q in zip(a + b, ga + gb):
                if q is None:
                    assert p is None
                else:
                    assert abs(p[0] - q[0]) < 1e-9 and abs(p[1] - q[1]) < 1e-9
        Z, H, W, C = heat.shape
        for z in range(Z):
            for c in range(C):
                ch = heat[z, :, :, c]
                if np.isnan(ch).any():
                    continue
                assert amax[z, c] == int(np.argmax(ch))
                assert vmax[z, c] == ch.max()


def test_angles_and_distances(gold):
    for key in gold['names']:
        lab = gold[key + '/labels']
        a, b = ex.ip_from_rvip_mask_3d(lab, keepdim=True, both_only=True)
        for p, q, ga, gd in zip(a, b, gold[key + '/angle'], gold[key + '/dist']):
            if p is None or q is None:
                assert np.isnan(ga) and np.isnan(gd)
            else:
                assert abs(ex.get_angle2x(p, q) - ga) < 1e-12
                assert abs(ex.get_dist(p, q) - gd) < 1e-12


def test_background_free_slice_quirk():
    """evaluate_cv.py:431 drops the smallest label when a slice has no background pixel."""
    m = np.ones((4, 4), np.uint8)
    m[2:, :] = 2
    a, b = ex.mean_rvip_2d(m, both_only=False)
    assert a is None and b == [2.5, 1.5]
