"""Where does model.predict(host volume) spend its time?  python profiles/predict_breakdown.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cmr_landmark_detection_b200.models.Unets import create_unet

torch.cuda.set_device(0)
m = create_unet(dict(bench.CONFIG))
x = np.random.default_rng(0).random((16, 256, 256, 1), dtype=np.float32)
for _ in range(5):
    m.predict(x, batch_size=16)
def t(f, n=30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print('predict(16)            %.3f ms' % t(lambda: m.predict(x, batch_size=16)))
print('predict(16, bs 8)      %.3f ms' % t(lambda: m.predict(x, batch_size=8)))
print('predict(16, bs 4)      %.3f ms' % t(lambda: m.predict(x, batch_size=4)))
pin = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
print('stage (1 thread copy)  %.3f ms' % t(lambda: np.copyto(pin.numpy(), x)))
print('stage (_stage)         %.3f ms' % t(lambda: m._stage(pin, x)))
xd = torch.empty(x.shape, dtype=torch.float32, device='cuda')
print('H2D pinned 4 MB        %.3f ms' % t(lambda: xd.copy_(pin, non_blocking=True)))
out = m.predict_device(xd)
print('forward device         %.3f ms' % t(lambda: m.predict_device(xd, out)))
po = torch.empty(out.shape, dtype=torch.float32, pin_memory=True)
print('D2H pinned 8 MB        %.3f ms' % t(lambda: po.copy_(out, non_blocking=True)))
print('pinned alloc 8 MB      %.3f ms' % t(lambda: torch.empty(out.shape, dtype=torch.float32, pin_memory=True)))
x1 = x[:1]
print('predict(1)             %.3f ms' % t(lambda: m.predict(x1, batch_size=1)))
xd1 = xd[:1].contiguous(); o1 = m.predict_device(xd1)
print('forward device (1)     %.3f ms' % t(lambda: m.predict_device(xd1, o1)))
