"""Times the conv kernels (forward epilogue mode) on the deep-level shapes of the bench network through the single-op
C-ABI entry points, and prints the halo kernel's issue-loop wait accounting (cycles, CTA 0).
usage: python profiles/conv_shapes.py [halo|tc] [reps]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cmr_landmark_detection_b200.runtime import ffi  # noqa: E402
from tests import gpu_util as U  # noqa: E402

SHAPES = [  # name, B, H, W, C0, C1, Cout
    ('mid.conv_b', 32, 16, 16, 512, 0, 512),
    ('mid.conv_a', 32, 16, 16, 256, 0, 512),
    ('dec0.upconv', 32, 32, 32, 512, 0, 256),
    ('dec0.conv_a', 32, 32, 32, 256, 256, 256),
    ('enc3.conv_b', 32, 32, 32, 256, 0, 256),
    ('dec1.upconv', 32, 64, 64, 256, 0, 128),
    ('dec1.up.dgrad', 32, 64, 64, 128, 0, 256),
    ('enc2.conv_b', 32, 64, 64, 128, 0, 128),
    ('dec2.upconv', 32, 128, 128, 128, 0, 64),
]
ROW_SHAPES = [
    ('enc0.conv_b', 32, 256, 256, 32, 0, 32),
    ('dec3.upconv', 32, 256, 256, 64, 0, 32),
    ('dec3.conv_a', 32, 256, 256, 32, 32, 32),
    ('dec3.up.dgrad', 32, 256, 256, 32, 0, 64),
    ('enc1.conv_a', 32, 128, 128, 32, 0, 64),
    ('enc1.conv_b', 32, 128, 128, 64, 0, 64),
    ('dec2.upconv', 32, 128, 128, 128, 0, 64),
    ('dec2.up.dgrad', 32, 128, 128, 64, 0, 128),
]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else 'halo'
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    fn = {'halo': 'rvip_conv3x3_halo', 'tc': 'rvip_conv3x3_tc', 'row': 'rvip_conv3x3_row'}[which]
    L = ffi.lib()
    g = torch.Generator(device='cuda').manual_seed(1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    dbg = torch.zeros((148, 8), dtype=torch.int64, device='cuda')
    if which in ('halo', 'row'):
        L.rvip_conv3x3_halo_debug(ffi.ptr(dbg))
    print('kernel,layer,us,TFLOP/s,loop_cycles,wait_tmem,wait_act,wait_weights,kernel_cycles,epilogue_cycles')
    for name, B, H, W, C0, C1, N in (ROW_SHAPES if which == 'row' else SHAPES):
        x0 = torch.randn((B, H, W, C0), generator=g, device='cuda').to(torch.bfloat16)
        x1 = torch.randn((B, H, W, C1), generator=g, device='cuda').to(torch.bfloat16) if C1 else None
        w = torch.randn((3, 3, C0 + C1, N), generator=g, device='cuda') * 0.02
        wp = U.pack_fwd(w)
        bias = torch.zeros(N, device='cuda')
        out = torch.empty((B, H, W, N), dtype=torch.bfloat16, device='cuda')
        stats = torch.zeros(2 * N, dtype=torch.float64, device='cuda')
        st = U.stream()

        def run():
            if which == 'row':
                ffi.check(L.rvip_conv3x3_row(ffi.ptr(x0), ffi.ptr(x1), C0, C1, ffi.ptr(wp), ffi.ptr(bias), ffi.ptr(out), None,
                                             N, ffi.ptr(stats), B, H, W, N, 0, 0, st))
            else:
                ffi.check(getattr(L, fn)(ffi.ptr(x0), ffi.ptr(x1), C0, C1, ffi.ptr(wp), ffi.ptr(bias), ffi.ptr(out), None,
                                         N, ffi.ptr(stats), B, H, W, N, 0, st))
        run()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        us = tot / reps * 1e3
        dbg.zero_()
        run()
        torch.cuda.synchronize()
        fl = 2.0 * 9 * (C0 + C1) * N * B * H * W
        d = dbg[0].tolist()
        print('%s,%s,%.1f,%.0f,%d,%d,%d,%d,%d,%d' % (which, name, us, fl / us / 1e6, d[0], d[1], d[2], d[3], d[4], d[5]))


if __name__ == '__main__':
    main()
