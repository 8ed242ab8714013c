"""Workload for `ncu`: W warm-up train steps + N profiled steps of the bench configuration (C2).
usage: python profiles/step_for_ncu.py [warmup] [steps] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import CONFIG  # noqa: E402
from cmr_landmark_detection_b200 import synth  # noqa: E402
from cmr_landmark_detection_b200.models.Unets import create_unet  # noqa: E402

warm = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
model = create_unet(dict(CONFIG))
x, y = synth.make_batch(B, 256, 256, seed=1)
xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
for i in range(warm + steps):
    model.train_step_device(xd, yd)
torch.cuda.synchronize()
print('launches per step', model.launch_count() // (warm + steps))
