"""Helpers shared by the GPU parity tests: call the single-op C-ABI entry points on torch tensors and
build torch fp32 references of the same ops (the oracle for the floating-point kernels)."""
import ctypes as C

import numpy as np
import torch
import torch.nn.functional as F

from cmr_landmark_detection_b200.runtime import ffi


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pack_fwd(w_hwio: torch.Tensor) -> torch.Tensor:
    """[3,3,C,N] -> Wf [N][9][C] bf16."""
    kh, kw, c, n = w_hwio.shape
    return w_hwio.permute(3, 0, 1, 2).reshape(n, 9 * c).contiguous().to(torch.bfloat16)


def pack_dgrad(w_hwio: torch.Tensor) -> torch.Tensor:
    """[3,3,C,N] -> Wd [C][9 (rotated)][N] bf16."""
    kh, kw, c, n = w_hwio.shape
    rot = torch.flip(w_hwio, dims=(0, 1))
    return rot.permute(2, 0, 1, 3).reshape(c, 9 * n).contiguous().to(torch.bfloat16)


def conv_tc(in0, in1, w_packed, bias, cout, mode, out_split=None, want_stats=False):
    B, H, W, C0 = in0.shape
    C1 = in1.shape[3] if in1 is not None else 0
    dev = in0.device
    if out_split is None:
        out_split = cout
    out0 = torch.full((B, H, W, out_split if mode == 2 else cout), float('nan'), dtype=torch.bfloat16, device=dev)
    out1 = None
    if mode == 2 and out_split < cout:
        out1 = torch.full((B, H, W, cout - out_split), float('nan'), dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=dev) if want_stats else None
    ffi.check(ffi.lib().rvip_conv3x3_tc(ffi.ptr(in0), ffi.ptr(in1), C0, C1, ffi.ptr(w_packed), ffi.ptr(bias),
                                        ffi.ptr(out0), ffi.ptr(out1), out_split, ffi.ptr(stats), B, H, W, cout, mode,
                                        stream()))
    torch.cuda.synchronize()
    return out0, out1, stats


def wgrad_tc(x0, x1, dz):
    B, H, W, C0 = x0.shape
    C1 = x1.shape[3] if x1 is not None else 0
    cout = dz.shape[3]
    dw = torch.zeros((3, 3, C0 + C1, cout), dtype=torch.float32, device=x0.device)
    ffi.check(ffi.lib().rvip_wgrad3x3_tc(ffi.ptr(x0), ffi.ptr(x1), C0, C1, ffi.ptr(dz), ffi.ptr(dw), B, H, W, cout,
                                         stream()))
    torch.cuda.synchronize()
    return dw


def ref_conv(x_nhwc: torch.Tensor, w_hwio: torch.Tensor, bias=None, relu=False):
    """fp32 reference on the bf16-rounded operands (what the tensor cores see), fp32 accumulate."""
    x = x_nhwc.float().permute(0, 3, 1, 2)
    w = w_hwio.to(torch.bfloat16).float().permute(3, 2, 0, 1)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        y = F.conv2d(x, w, bias, padding=1)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_err(a, b) -> float:
    return float((a.double() - b.double()).abs().max())


def conv_row(in0, in1, w_packed, bias, cout, mode, out_split=None, want_stats=False, base_offset_mode=0):
    B, H, W, C0 = in0.shape
    C1 = in1.shape[3] if in1 is not None else 0
    dev = in0.device
    if out_split is None:
        out_split = cout
    out0 = torch.full((B, H, W, out_split if mode == 2 else cout), float('nan'), dtype=torch.bfloat16, device=dev)
    out1 = None
    if mode == 2 and out_split < cout:
        out1 = torch.full((B, H, W, cout - out_split), float('nan'), dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=dev) if want_stats else None
    ffi.check(ffi.lib().rvip_conv3x3_row(ffi.ptr(in0), ffi.ptr(in1), C0, C1, ffi.ptr(w_packed), ffi.ptr(bias),
                                         ffi.ptr(out0), ffi.ptr(out1), out_split, ffi.ptr(stats), B, H, W, cout, mode,
                                         base_offset_mode, stream()))
    torch.cuda.synchronize()
    return out0, out1, stats


def wgrad_row(x0, x1, dz):
    B, H, W, C0 = x0.shape
    C1 = x1.shape[3] if x1 is not None else 0
    cout = dz.shape[3]
    dw = torch.zeros((3, 3, C0 + C1, cout), dtype=torch.float32, device=x0.device)
    ffi.check(ffi.lib().rvip_wgrad3x3_row(ffi.ptr(x0), ffi.ptr(x1), C0, C1, ffi.ptr(dz), ffi.ptr(dw), B, H, W, cout,
                                          stream()))
    torch.cuda.synchronize()
    return dw


def wgrad_halo(x0, x1, dz):
    B, H, W, C0 = x0.shape
    C1 = x1.shape[3] if x1 is not None else 0
    cout = dz.shape[3]
    dw = torch.zeros((3, 3, C0 + C1, cout), dtype=torch.float32, device=x0.device)
    ffi.check(ffi.lib().rvip_wgrad3x3_halo(ffi.ptr(x0), ffi.ptr(x1), C0, C1, ffi.ptr(dz), ffi.ptr(dw), B, H, W, cout,
                                           stream()))
    torch.cuda.synchronize()
    return dw


def conv_halo(in0, in1, w_packed, bias, cout, mode, out_split=None, want_stats=False):
    B, H, W, C0 = in0.shape
    C1 = in1.shape[3] if in1 is not None else 0
    dev = in0.device
    if out_split is None:
        out_split = cout
    out0 = torch.full((B, H, W, out_split if mode == 2 else cout), float('nan'), dtype=torch.bfloat16, device=dev)
    out1 = None
    if mode == 2 and out_split < cout:
        out1 = torch.full((B, H, W, cout - out_split), float('nan'), dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=dev) if want_stats else None
    ffi.check(ffi.lib().rvip_conv3x3_halo(ffi.ptr(in0), ffi.ptr(in1), C0, C1, ffi.ptr(w_packed), ffi.ptr(bias),
                                          ffi.ptr(out0), ffi.ptr(out1), out_split, ffi.ptr(stats), B, H, W, cout, mode,
                                          stream()))
    torch.cuda.synchronize()
    return out0, out1, stats


def upconv_halo(direction, x_or_dz, w_hwio, bias=None, transposed=False):
    """Phase-decomposed up-convolution through the C ABI.  direction 0: x [B,h,w,Cin] -> u [B,2h,2w,C];
    direction 1: dz [B,2h,2w,C] -> dx [B,h,w,Cin].  transposed: w is a Conv2DTranspose kernel (kh, kw, C, Cin)."""
    cin, c = (w_hwio.shape[3], w_hwio.shape[2]) if transposed else (w_hwio.shape[2], w_hwio.shape[3])
    dev = x_or_dz.device
    if direction == 0:
        B, h, w, _ = x_or_dz.shape
        low = x_or_dz
        high = torch.full((B, 2 * h, 2 * w, c), float('nan'), dtype=torch.bfloat16, device=dev)
        out = high
    else:
        B, H, W, _ = x_or_dz.shape
        h, w = H // 2, W // 2
        high = x_or_dz
        low = torch.full((B, h, w, cin), float('nan'), dtype=torch.bfloat16, device=dev)
        out = low
    scratch = torch.zeros(2 * max(16 * cin * c, 768 * cin), dtype=torch.bfloat16, device=dev)
    wf = w_hwio.float().contiguous()
    ffi.check(ffi.lib().rvip_upconv3x3_halo(direction, ffi.ptr(low), ffi.ptr(high), ffi.ptr(wf), ffi.ptr(bias),
                                            ffi.ptr(scratch), B, h, w, cin, c, int(transposed), stream()))
    torch.cuda.synchronize()
    return out


def ref_upconv(x_low_nhwc, w_hwio, bias=None, relu=True, dz=None):
    """fp32 reference: nearest x2 up-sampling followed by the 3x3 convolution (weights NOT pre-summed), and
    optionally the gradient w.r.t. the low-resolution input for an output gradient dz."""
    x = x_low_nhwc.float().permute(0, 3, 1, 2).clone().requires_grad_(dz is not None)
    w = w_hwio.float().permute(3, 2, 0, 1)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        y = F.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), w, bias, padding=1)
    if dz is not None:
        y.backward(dz.float().permute(0, 3, 1, 2))
        return x.grad.permute(0, 2, 3, 1).contiguous()
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


def upconv_wgrad_halo(x_low, dz, transposed=False):
    """Weight gradient of the phase-decomposed up-convolution: x [B,h,w,Cin], dz [B,2h,2w,C] -> dw [3,3,Cin,C] fp32
    (transposed: the Conv2DTranspose kernel gradient [3,3,C,Cin])."""
    B, h, w, cin = x_low.shape
    c = dz.shape[3]
    dw = torch.zeros((3, 3, c, cin) if transposed else (3, 3, cin, c), dtype=torch.float32, device=x_low.device)
    ffi.check(ffi.lib().rvip_upconv_wgrad_halo(ffi.ptr(x_low), ffi.ptr(dz), ffi.ptr(dw), B, h, w, cin, c, int(transposed),
                                               stream()))
    torch.cuda.synchronize()
    return dw


def ref_tconv(x_low_nhwc, w_hwoi, bias=None, relu=True):
    """fp32 reference of Conv2DTranspose(3, strides 2, 'same') [-> ReLU] on bf16-rounded weights; returns (y NHWC, the
    autograd leaves (x, w)) so that callers can pull gradients."""
    x = x_low_nhwc.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    w = w_hwoi.to(torch.bfloat16).float().clone().requires_grad_(True)
    h, wd = x.shape[2], x.shape[3]
    y = F.conv_transpose2d(x, w.permute(3, 2, 0, 1), bias, stride=2)[:, :, :2 * h, :2 * wd]
    if relu:
        y = torch.relu(y)
    return y, x, w
