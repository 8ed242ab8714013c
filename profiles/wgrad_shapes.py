"""Times (CUDA events, 20 reps, L2 flushed between reps by a 256 MB memset) the wgrad kernels on the deep-level
shapes of the bench network, through the single-op C-ABI entry points.  Also the workload for ncu captures.
usage: python profiles/wgrad_shapes.py [halo|tc|row] [reps]"""
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
    ('enc3.conv_a', 32, 32, 32, 128, 0, 256),
    ('dec1.upconv', 32, 64, 64, 256, 0, 128),
    ('enc2.conv_a', 32, 64, 64, 64, 0, 128),
    ('dec2.upconv', 32, 128, 128, 128, 0, 64),
    ('enc1.conv_b', 32, 128, 128, 64, 0, 64),
    ('dec3.upconv', 32, 256, 256, 64, 0, 32),
]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else 'halo'
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    fn = {'halo': 'rvip_wgrad3x3_halo', 'tc': 'rvip_wgrad3x3_tc', 'row': 'rvip_wgrad3x3_row'}[which]
    L = ffi.lib()
    g = torch.Generator(device='cuda').manual_seed(1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    print('kernel,layer,us,TFLOP/s')
    for name, B, H, W, C0, C1, N in SHAPES:
        if which == 'row' and W % 128:
            continue
        x0 = torch.randn((B, H, W, C0), generator=g, device='cuda').to(torch.bfloat16)
        x1 = torch.randn((B, H, W, C1), generator=g, device='cuda').to(torch.bfloat16) if C1 else None
        dz = torch.randn((B, H, W, N), generator=g, device='cuda').to(torch.bfloat16)
        dw = torch.zeros((3, 3, C0 + C1, N), dtype=torch.float32, device='cuda')
        st = U.stream()

        def run():
            ffi.check(getattr(L, fn)(ffi.ptr(x0), ffi.ptr(x1), C0, C1, ffi.ptr(dz), ffi.ptr(dw), B, H, W, N, st))
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
        fl = 2.0 * 9 * (C0 + C1) * N * B * H * W
        print('%s,%s,%.1f,%.0f' % (which, name, us, fl / us / 1e6))


if __name__ == '__main__':
    main()
