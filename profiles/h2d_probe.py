"""H2D / D2H copy bandwidth from page-locked host memory of different flavours (one GPU).  python profiles/h2d_probe.py"""
import ctypes as C, time, sys
import numpy as np, torch
torch.cuda.set_device(0)
rt = C.CDLL('libcudart.so.12') if True else None
N = 25 * 1024 * 1024
dev = torch.empty(N, dtype=torch.uint8, device='cuda')
def bw(host_ptr, label, d2h=False, reps=20):
    s = torch.cuda.current_stream().cuda_stream
    kind = 2 if d2h else 1
    for _ in range(3):
        rt.cudaMemcpyAsync(C.c_void_p(host_ptr if d2h else dev.data_ptr()), C.c_void_p(dev.data_ptr() if d2h else host_ptr), C.c_size_t(N), kind, C.c_void_p(s))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        rt.cudaMemcpyAsync(C.c_void_p(host_ptr if d2h else dev.data_ptr()), C.c_void_p(dev.data_ptr() if d2h else host_ptr), C.c_size_t(N), kind, C.c_void_p(s))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('%-34s %s %.3f ms  %.1f GB/s' % (label, 'D2H' if d2h else 'H2D', ms, N / ms / 1e6), flush=True)
t = torch.empty(N, dtype=torch.uint8, pin_memory=True); t.fill_(1)
bw(t.data_ptr(), 'torch pin_memory'); bw(t.data_ptr(), 'torch pin_memory', True)
for flags, name in ((0, 'cudaHostAlloc default'), (1, 'cudaHostAlloc portable'), (4, 'cudaHostAlloc write-combined')):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), flags) == 0
    C.memset(p, 1, N)
    bw(p.value, name); bw(p.value, name, True)
a = np.ones(N + 4096, dtype=np.uint8)
addr = (a.ctypes.data + 4095) & ~4095
assert rt.cudaHostRegister(C.c_void_p(addr), C.c_size_t(N), 0) == 0
bw(addr, 'cudaHostRegister(numpy)'); bw(addr, 'cudaHostRegister(numpy)', True)
t0 = time.perf_counter(); src = np.ones(N, dtype=np.uint8)
for flags, name in ((0, 'default'), (4, 'write-combined')):
    p = C.c_void_p(); rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), flags)
    dst = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(N,))
    np.copyto(dst, src)
    t0 = time.perf_counter()
    for _ in range(10): np.copyto(dst, src)
    print('host memcpy into %-16s %.3f ms (25 MiB, 1 thread)' % (name, (time.perf_counter() - t0) / 10 * 1e3))
