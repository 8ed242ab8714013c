import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from cmr_landmark_detection_b200 import synth
from cmr_landmark_detection_b200.models.Unets import create_unet
from cmr_landmark_detection_b200.extract import extract_device
model = create_unet(dict(bench.CONFIG))
dev = torch.device('cuda', 0)
for B in (16, 32, 64, 128):
    x, _ = synth.make_batch(B, 256, 256, seed=3)
    xd = torch.from_numpy(x).to(dev)
    for _ in range(3):
        extract_device(model.predict_device(xd))
    torch.cuda.synchronize()
    n = 20
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = extract_device(model.predict_device(xd))
    e1.record()
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print('B=%d: %.3f ms per call (cpu enqueue %.3f ms), %.0f vols/s, %.0f slices/s' % (B, ms, t_cpu / n * 1e3, B / 16 / ms * 1e3, B / ms * 1e3))
model.profile(16, False, True)
x, _ = synth.make_batch(16, 256, 256, seed=3)
xd = torch.from_numpy(x).to(dev)
for _ in range(10):
    model.predict_device(xd)
torch.cuda.synchronize()
print({k: (round(v[0] / 10 * 1e3, 1), v[1] // 10) for k, v in model.profile_read(16, False).items() if v[1]})
