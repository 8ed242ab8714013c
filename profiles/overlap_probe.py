"""How well does a tensor-pipe-bound wgrad kernel share the GPU with an HBM-bound streaming kernel on a second stream?
Times the halo wgrad (C-ABI single-op entry) alone, a torch element-wise pass over a 134 MB bf16 tensor alone, and both
launched concurrently on two streams.  usage: python profiles/overlap_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402

import torch  # noqa: E402

from cmr_landmark_detection_b200.runtime import ffi  # noqa: E402

L = ffi.lib()
g = torch.Generator(device='cuda').manual_seed(1)
SH = [('dec3.upconv 64->32 @256', 32, 256, 256, 64, 32), ('dec1.upconv 256->128 @64', 32, 64, 64, 256, 128)]
a = torch.randn((32, 256, 256, 32), generator=g, device='cuda').to(torch.bfloat16)     # 134 MB
b = torch.empty_like(a)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print('case,wgrad_alone_us,stream_alone_us,both_us,sum_us')
for name, B, H, W, Ci, Co in SH:
    x0 = torch.randn((B, H, W, Ci), generator=g, device='cuda').to(torch.bfloat16)
    dz = torch.randn((B, H, W, Co), generator=g, device='cuda').to(torch.bfloat16)
    dw = torch.zeros((3, 3, Ci, Co), dtype=torch.float32, device='cuda')

    def wgrad(stream):
        ffi.check(L.rvip_wgrad3x3_halo(ffi.ptr(x0), None, Ci, 0, ffi.ptr(dz), ffi.ptr(dw), B, H, W, Co,
                                       C.c_void_p(stream.cuda_stream)))

    def stream_op(stream):
        with torch.cuda.stream(stream):
            torch.mul(a, 2.0, out=b)

    def both():
        ev = torch.cuda.Event()
        ev.record()
        s1.wait_event(ev)
        s2.wait_event(ev)
        wgrad(s1)
        stream_op(s2)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
    t1 = timeit(lambda: (wgrad(s1), torch.cuda.current_stream().wait_stream(s1)))
    t2 = timeit(lambda: (stream_op(s2), torch.cuda.current_stream().wait_stream(s2)))
    t12 = timeit(both)
    print('%s,%.1f,%.1f,%.1f,%.1f' % (name, t1, t2, t12, t1 + t2))
