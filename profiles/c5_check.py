"""Config C5 (BASELINE.json): 5-level U-Net, 64 base filters, 512x512, batch 8 -- runs a few training steps, prints the
step time, conv TFLOP/s and which kernel each layer class used.  usage: python profiles/c5_check.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from cmr_landmark_detection_b200 import synth  # noqa: E402
from cmr_landmark_detection_b200.models.Unets import create_unet  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cfg = dict(bench.CONFIG, DIM=[512, 512], DEPTH=5, FILTERS=64)
model = create_unet(cfg)
print('params', model.count_params())
B = 8
x, y = synth.make_batch(B, 512, 512, seed=1)
xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
losses = []
for _ in range(3):
    losses.append(float(model.train_step_device(xd, yd).item()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = model.train_step_device(xd, yd)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
fl = bench.conv_flops_per_slice(cfg)
print('C5: %.2f ms/step, %.1f slices/s, %.0f TFLOP/s (conv, fwd+dgrad+wgrad), losses %s -> %.5f' %
      (ms, B / ms * 1e3, fl['train'] * B / ms / 1e9, ['%.5f' % l for l in losses], float(loss.item())))
assert np.isfinite(float(loss.item()))
heat = model.predict(x[:2], batch_size=2)
print('predict', heat.shape, float(heat.mean()))
print('peak memory GB', torch.cuda.max_memory_allocated() / 1e9)
