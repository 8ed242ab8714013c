"""Device-backed stand-in for src/data/Postprocess.py:108-120 (same name, argument and return structure)."""
from __future__ import annotations

import numpy as np
import torch

from ..extract import cc_filter_device


def clean_3d_prediction_2d_cc(pred):
    """[Z,H,W] label volume -> same shape and dtype, only the biggest connected component of each label per slice."""
    pred = np.asarray(pred)
    assert pred.ndim == 3, 'invalid shape: {}'.format(pred.shape)
    dev = torch.device('cuda', torch.cuda.current_device())
    out = cc_filter_device(torch.from_numpy(np.ascontiguousarray(pred.astype(np.uint8))).to(dev), 8)
    return out.cpu().numpy().astype(pred.dtype)
