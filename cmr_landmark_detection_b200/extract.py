"""Heat map -> landmarks on the device.  Replaces the host loops of
src/models/predict_model.py:149-156 (threshold -> label map) and src/models/evaluate_cv.py:389-442
(get_ip_from_rvip_mask_3d / get_mean_rvip_2d) and adds the argmax/max of SURVEY row E3.  The return
structures are the reference's: two lists (anterior, inferior) of [y, x] float64 points or None."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .runtime import ffi


def extract_device(heat: torch.Tensor, thr: float = 0.5) -> Dict[str, torch.Tensor]:
    """heat [Z,H,W,C] fp32 CUDA tensor -> device tensors: yx [Z,C,2] f64 (NaN = label absent),
    count [Z,C] i32, argmax [Z,C] i32 (flat row-major index), maxv [Z,C] f32."""
    if not heat.is_cuda or heat.dtype != torch.float32 or heat.dim() != 4:
        raise ValueError('extract_device expects a CUDA float32 [Z,H,W,C] tensor')
    heat = heat.contiguous()
    Z, H, W, Cc = heat.shape
    dev = heat.device
    L = ffi.lib()
    yx = torch.empty((Z, Cc, 2), dtype=torch.float64, device=dev)
    count = torch.empty((Z, Cc), dtype=torch.int32, device=dev)
    argmax = torch.empty((Z, Cc), dtype=torch.int32, device=dev)
    maxv = torch.empty((Z, Cc), dtype=torch.float32, device=dev)
    scratch = torch.empty(max(int(L.rvip_extract_scratch_bytes(Z, Cc)), 8), dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ffi.check(L.rvip_extract(ffi.ptr(heat), Z, H, W, Cc, float(thr), ffi.ptr(yx), ffi.ptr(count), ffi.ptr(argmax),
                             ffi.ptr(maxv), ffi.ptr(scratch), st))
    return {'yx': yx, 'count': count, 'argmax': argmax, 'maxv': maxv}


def label_map_device(heat: torch.Tensor, thr: float = 0.5) -> torch.Tensor:
    """predict_model.py:153-156 on the device: [.., C] fp32 -> [..] uint8."""
    heat = heat.contiguous()
    out = torch.empty(heat.shape[:-1], dtype=torch.uint8, device=heat.device)
    st = C.c_void_p(torch.cuda.current_stream(heat.device).cuda_stream)
    ffi.check(ffi.lib().rvip_label_map(ffi.ptr(heat), out.numel(), heat.shape[-1], float(thr), ffi.ptr(out), st))
    return out


def cc_filter_device(labels: torch.Tensor, connectivity: int = 8) -> torch.Tensor:
    """Largest-connected-component filter on the device: labels [Z,H,W] uint8 CUDA tensor -> same shape, only the
    largest component of each label value per slice (Postprocess.py:108-120; 8-connected like the reference's
    cv2 call actually runs)."""
    if not labels.is_cuda or labels.dtype != torch.uint8 or labels.dim() != 3:
        raise ValueError('cc_filter_device expects a CUDA uint8 [Z,H,W] tensor')
    labels = labels.contiguous()
    Z, H, W = labels.shape
    L = ffi.lib()
    out = torch.empty_like(labels)
    scratch = torch.empty(int(L.rvip_cc_scratch_bytes(Z, H, W)) // 8 + 1, dtype=torch.int64, device=labels.device)
    st = C.c_void_p(torch.cuda.current_stream(labels.device).cuda_stream)
    ffi.check(L.rvip_cc_filter(ffi.ptr(labels), Z, H, W, int(connectivity), ffi.ptr(out), ffi.ptr(scratch), st))
    return out


def points_from_stats(yx: np.ndarray, count: np.ndarray, keepdim: bool = False, both_only: bool = True):
    """(yx, count) -> the reference's (first_ips, second_ips) lists (evaluate_cv.py:389-442)."""
    first, second = [], []
    for z in range(yx.shape[0]):
        pts = [None if count[z, c] == 0 else [float(yx[z, c, 0]), float(yx[z, c, 1])] for c in range(2)]
        if both_only and (pts[0] is None or pts[1] is None):
            pts = [None, None]
        if (pts[0] is not None and pts[1] is not None) or keepdim:
            first.append(pts[0])
            second.append(pts[1])
    return first, second


def get_ip_from_heatmaps(preds, thr: float = 0.5, keepdim: bool = False, both_only: bool = True, device=None,
                         cc_filter: bool = False):
    """model.predict output [Z,H,W,2] (ndarray or CUDA tensor) -> (anterior list, inferior list):
    == get_ip_from_rvip_mask_3d(label_map(preds)) of the reference, fused on the device.  cc_filter=True inserts the
    largest-connected-component filter of predict_model.py:159-161 (CC_FILTER) between the two."""
    if isinstance(preds, np.ndarray):
        dev = device or torch.device('cuda', torch.cuda.current_device())
        preds = torch.from_numpy(np.ascontiguousarray(preds, dtype=np.float32)).to(dev)
    if cc_filter:
        lab = cc_filter_device(label_map_device(preds, thr))
        preds = torch.stack([lab == c + 1 for c in range(preds.shape[-1])], dim=-1).to(torch.float32)
        thr = 0.5
    r = extract_device(preds, thr)
    return points_from_stats(r['yx'].cpu().numpy(), r['count'].cpu().numpy(), keepdim=keepdim, both_only=both_only)


def landmark_metrics_device(gt_yx: torch.Tensor, pred_yx: torch.Tensor, spacing: float = 1.0, threshold: float = 1000.0,
                            dim: float = 224.0) -> Dict[str, torch.Tensor]:
    """Per-volume landmark metrics on the device (evaluate_cv.py: get_angle2x, get_distances,
    get_distances_upper_bound, calc_mean_ip, calc_tpr_thresh, calc_ppv_thresh).  gt_yx / pred_yx: CUDA float64
    [Z, 2, 2] = [slice][anterior, inferior][y, x] with NaN for a missing point (extract_device's 'yx' layout)."""
    if gt_yx.shape != pred_yx.shape or gt_yx.dim() != 3 or gt_yx.shape[1:] != (2, 2):
        raise ValueError('landmark_metrics_device expects two [Z, 2, 2] tensors')
    gt_yx = gt_yx.to(torch.float64).contiguous()
    pred_yx = pred_yx.to(torch.float64).contiguous()
    Z, dev = gt_yx.shape[0], gt_yx.device
    angle = torch.empty((2, Z), dtype=torch.float64, device=dev)
    dist = torch.empty((2, Z), dtype=torch.float64, device=dev)
    dist_thr = torch.empty((2, Z), dtype=torch.float64, device=dev)
    dist_ub = torch.empty((2, Z), dtype=torch.float64, device=dev)
    summary = torch.empty(18, dtype=torch.float64, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ffi.check(ffi.lib().rvip_landmark_metrics(ffi.ptr(gt_yx), ffi.ptr(pred_yx), Z, float(spacing), float(threshold),
                                              float(dim), ffi.ptr(angle), ffi.ptr(dist), ffi.ptr(dist_thr),
                                              ffi.ptr(dist_ub), ffi.ptr(summary), st))
    return {'angle_gt': angle[0], 'angle_pred': angle[1], 'dist': dist, 'dist_thr': dist_thr, 'dist_ub': dist_ub,
            'mean_ip': summary[:8].view(2, 2, 2), 'tpr': summary[8:10], 'ppv': summary[10:12],
            'counters': summary[12:18].view(2, 3)}


def _points_to_tensor(ips, device) -> torch.Tensor:
    """(anterior list, inferior list) of [y, x] / None -> float64 [Z, 2, 2] with NaN rows."""
    ants, infs = ips
    a = np.array([[np.nan, np.nan] if p is None else [float(p[0]), float(p[1])] for p in ants], np.float64).reshape(-1, 2)
    b = np.array([[np.nan, np.nan] if p is None else [float(p[0]), float(p[1])] for p in infs], np.float64).reshape(-1, 2)
    return torch.from_numpy(np.stack([a, b], axis=1)).to(device)
