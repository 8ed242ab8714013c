"""Issue-loop wait accounting of the phase-decomposed up-convolution kernels (conv_halo.cu, NS = 2 / 3) on the four
decoder up-conv shapes of the bench network, forward and dgrad, through rvip_upconv3x3_halo (which also packs the
weights, so only the in-kernel cycle counters are meaningful here, not wall time).
usage: python profiles/upconv_shapes.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cmr_landmark_detection_b200.runtime import ffi  # noqa: E402
from tests import gpu_util as U  # noqa: E402

SHAPES = [('dec0.upconv', 32, 16, 16, 512, 256), ('dec1.upconv', 32, 32, 32, 256, 128),
          ('dec2.upconv', 32, 64, 64, 128, 64), ('dec3.upconv', 32, 128, 128, 64, 32)]


def main():
    L = ffi.lib()
    g = torch.Generator(device='cuda').manual_seed(1)
    dbg = torch.zeros((148, 8), dtype=torch.int64, device='cuda')
    L.rvip_conv3x3_halo_debug(ffi.ptr(dbg))
    print('layer,dir,ctas,kernel_cycles,kernel_us_at_1965MHz,loop_cycles,wait_tmem,wait_act,wait_weights,epilogue_cycles')
    for name, B, h, w, cin, c in SHAPES:
        wt = torch.randn((3, 3, cin, c), generator=g, device='cuda') * 0.02
        for d in (0, 1):
            t = torch.randn((B, h, w, cin) if d == 0 else (B, 2 * h, 2 * w, c), generator=g, device='cuda').to(torch.bfloat16)
            U.upconv_halo(d, t, wt, torch.zeros(c, device='cuda') if d == 0 else None)
            dbg.zero_()
            U.upconv_halo(d, t, wt, torch.zeros(c, device='cuda') if d == 0 else None)
            v = dbg.cpu().double()
            used = v[:, 4] > 0
            m = v[used].mean(dim=0)
            print('%s,%s,%d,%.0f,%.1f,%.0f,%.0f,%.0f,%.0f,%.0f' % (name, 'fwd' if d == 0 else 'dgrad', int(used.sum()), m[4],
                                                                 m[4] / 1965.0, m[0], m[1], m[2], m[3], m[5]))
    L.rvip_conv3x3_halo_debug(None)


if __name__ == '__main__':
    main()
