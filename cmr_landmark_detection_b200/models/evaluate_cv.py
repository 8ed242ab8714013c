"""Device-backed stand-ins for the landmark functions of src/models/evaluate_cv.py (same names,
arguments and return structures); see ..extract for the kernel entry."""
from __future__ import annotations

import numpy as np
import torch

from ..extract import extract_device, get_ip_from_heatmaps, points_from_stats  # noqa: F401


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


def get_mean_rvip_2d(nda_2d, both_only=False):
    """evaluate_cv.py:418-442."""
    nda_2d = np.asarray(nda_2d)
    assert len(nda_2d.shape) == 2, 'invalid shape: {}'.format(nda_2d.shape)
    a, b = get_ip_from_rvip_mask_3d(nda_2d[None], keepdim=True, both_only=both_only)
    return a[0], b[0]
