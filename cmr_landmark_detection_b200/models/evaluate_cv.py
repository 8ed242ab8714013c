"""Device-backed stand-ins for the landmark functions of src/models/evaluate_cv.py (same names,
arguments and return structures); see ..extract for the kernel entry."""
from __future__ import annotations

import numpy as np
import torch

from ..extract import (extract_device, get_ip_from_heatmaps, landmark_metrics_device, points_from_stats,  # noqa: F401
                       _points_to_tensor)


def _onehot(msk: np.ndarray) -> torch.Tensor:
    dev = torch.device('cuda', torch.cuda.current_device())
    m = torch.from_numpy(np.ascontiguousarray(msk)).to(dev)
    return torch.stack([(m == 1), (m == 2)], dim=-1).to(torch.float32)


def get_ip_from_rvip_mask_3d(msk_3d, debug=False, keepdim=False, both_only=True):
    """evaluate_cv.py:389-416: label volume [Z,H,W] (0 / 1 anterior / 2 inferior) -> two lists of [y, x]."""
    msk_3d = np.asarray(msk_3d)
    assert msk_3d.ndim == 3, 'invalid shape: {}'.format(msk_3d.shape)
    r = extract_device(_onehot(msk_3d), 0.5)
    return points_from_stats(r['yx'].cpu().numpy(), r['count'].cpu().numpy(), keepdim=keepdim, both_only=both_only)


def get_ip_from_rvip_file(f_name, keepdim=False, both_only=True):
    """evaluate_cv.py:385-387: label volume file -> insertion points.  The reference reads through SimpleITK; the .nrrd
    volumes pred_fold writes here are read by utils/nrrd_io.py."""
    from ..utils.nrrd_io import read_nrrd
    vol, _ = read_nrrd(f_name)
    return get_ip_from_rvip_mask_3d(vol, keepdim=keepdim, both_only=both_only)


def get_mean_rvip_2d(nda_2d, both_only=False):
    """evaluate_cv.py:418-442."""
    nda_2d = np.asarray(nda_2d)
    assert len(nda_2d.shape) == 2, 'invalid shape: {}'.format(nda_2d.shape)
    a, b = get_ip_from_rvip_mask_3d(nda_2d[None], keepdim=True, both_only=both_only)
    return a[0], b[0]


def _nan_to_none(t):
    return np.array([None if not np.isfinite(v) else float(v) for v in t.cpu().numpy()], dtype=object)


def _metrics(ips1, ips2, **kw):
    dev = torch.device('cuda', torch.cuda.current_device())
    return landmark_metrics_device(_points_to_tensor(ips1, dev), _points_to_tensor(ips2, dev), **kw)


def get_distances(ips1, ips2, spacing=1, threshold=None):
    """evaluate_cv.py:549-561 on the device: (anterior distances, inferior distances), None where undefined."""
    m = _metrics(ips1, ips2, spacing=spacing, threshold=1e300 if threshold is None else threshold)
    d = m['dist'] if threshold is None else m['dist_thr']
    return _nan_to_none(d[0]), _nan_to_none(d[1])


def get_distances_upper_bound(ips1, ips2, spacing=1, dim=224):
    """evaluate_cv.py:572-595 on the device (ips1 = ground truth, ips2 = prediction)."""
    m = _metrics(ips1, ips2, spacing=spacing, dim=dim)
    return _nan_to_none(m['dist_ub'][0]), _nan_to_none(m['dist_ub'][1])


def calc_tpr_thresh(gt, pred, thresh=1000, spacing=1):
    """evaluate_cv.py:267-308 on the device."""
    t = _metrics(gt, pred, spacing=spacing, threshold=thresh)['tpr'].cpu().numpy()
    return float(t[0]), float(t[1])


def calc_ppv_thresh(gt, pred, thresh=1000, spacing=1):
    """evaluate_cv.py:311-353 on the device."""
    t = _metrics(gt, pred, spacing=spacing, threshold=thresh)['ppv'].cpu().numpy()
    return float(t[0]), float(t[1])


def calc_mean_ip(ips_list):
    """evaluate_cv.py:113-120 on the device: (mean anterior, mean inferior), NaN unless both lists have points."""
    m = _metrics(ips_list, ips_list)['mean_ip'][0].cpu().numpy()
    if not np.isfinite(m).all():
        return np.nan, np.nan
    return m[0], m[1]
