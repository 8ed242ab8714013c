"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's per-volume landmark metrics, NaN-coded.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(cmr_landmark_detection_b200/) never does.  Pinned: tests/golden/metrics_golden.npz holds outputs of the reference's
own functions (imported in the build container by tests/golden/make_metrics_golden.py).

Points are float64 [Z, 2] arrays of (y, x); a missing point (the reference's None) is a NaN row.
Reference functions restated (src/models/evaluate_cv.py):
  get_angle2x              :508-536   angle of the anterior->inferior line against the x axis, degrees in [0, 360)
  get_dist / get_distances :538-561   Euclidean distance * spacing, None when a point is missing, optional threshold
  get_distances_upper_bound:572-595   missing prediction -> distance to the farthest image corner (dim x dim)
  calc_mean_ip             :113-120   mean anterior / inferior point over the slices that have them
  calc_tpr_thresh          :267-308   per-landmark TPR with a distance threshold
  calc_ppv_thresh          :311-353   per-landmark PPV with a distance threshold
"""
from math import atan2, degrees

import numpy as np


def _present(p):
    return np.isfinite(p).all(axis=-1)


def angles(ant, inf):
    """[Z,2],[Z,2] -> [Z] degrees (NaN where a point is missing)  -- get_angle2x per slice."""
    out = np.full(len(ant), np.nan)
    for z in range(len(ant)):
        if _present(ant[z]) and _present(inf[z]):
            a = degrees(atan2(inf[z, 0] - ant[z, 0], inf[z, 1] - ant[z, 1]))
            out[z] = 360 + a if a < 0 else a
    return out


def distances(p1, p2, spacing=1.0, threshold=None):
    """get_distances for one landmark: [Z] distances, NaN when missing (or above the threshold)."""
    out = np.full(len(p1), np.nan)
    for z in range(len(p1)):
        if _present(p1[z]) and _present(p2[z]):
            d = float(np.linalg.norm(p1[z] - p2[z])) * spacing
            if threshold is None or d <= threshold:
                out[z] = d
    return out


def distances_upper_bound(gt, pred, spacing=1.0, dim=224):
    """get_distances_upper_bound for one landmark."""
    out = np.full(len(gt), np.nan)
    corners = np.array([(0, 0), (0, dim), (dim, 0), (dim, dim)], np.float64)
    for z in range(len(gt)):
        if _present(gt[z]) and _present(pred[z]):
            out[z] = float(np.linalg.norm(gt[z] - pred[z])) * spacing
        elif _present(gt[z]):
            out[z] = max(float(np.linalg.norm(gt[z] - c)) * spacing for c in corners)
    return out


def mean_ip(ant, inf):
    """calc_mean_ip: (mean anterior [2], mean inferior [2]); both NaN unless each list has at least one point."""
    pa, pi = _present(ant), _present(inf)
    if pa.any() and pi.any():
        return ant[pa].mean(axis=0), inf[pi].mean(axis=0)
    return np.full(2, np.nan), np.full(2, np.nan)


def tpr_ppv(gt, pred, thresh=1000.0, spacing=1.0):
    """calc_tpr_thresh / calc_ppv_thresh for one landmark -> (tpr, ppv, tp, fn, fp)."""
    tp = fn = fp = 0
    for z in range(len(gt)):
        g, p = _present(gt[z]), _present(pred[z])
        if g and p:
            if float(np.linalg.norm(gt[z] - pred[z])) * spacing <= thresh:
                tp += 1
            else:
                fp += 1
        elif g:
            fn += 1
        elif p:
            fp += 1
    tpr = tp / (tp + fn) if tp > 0 else 0
    ppv = tp / (tp + fp) if tp > 0 else 0
    return tpr, ppv, tp, fn, fp
