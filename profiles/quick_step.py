"""Quick A/B probe: overlapped ms/step (CUDA events, 30 steps) + per-class profile-mode times of the bench step for the
current environment.  Usage: [ENV=...] python profiles/quick_step.py [label]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from cmr_landmark_detection_b200 import synth  # noqa: E402
from cmr_landmark_detection_b200.models.Unets import create_unet  # noqa: E402


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else ''
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    model = create_unet(dict(bench.CONFIG))
    B = bench.BATCH_PER_GPU
    data = [synth.make_batch(B, 256, 256, seed=42 + i) for i in range(2)]
    dd = [(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)) for x, y in data]
    for i in range(5):
        model.train_step_device(*dd[i % 2])
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            loss = model.train_step_device(*dd[i % 2])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    model.profile(B, True, True)
    for i in range(10):
        model.train_step_device(*dd[i % 2])
    torch.cuda.synchronize()
    prof = model.profile_read(B, True)
    model.profile(B, True, False)
    cls = {k: round(v[0] / 10, 4) for k, v in prof.items() if v[0] > 0}
    print(json.dumps({'label': label, 'ms_per_step': round(best, 4), 'slices_per_s': round(B / best * 1e3, 1),
                      'loss': float(loss.item()), 'profile_sum_ms': round(sum(cls.values()), 4), 'classes': cls}))


if __name__ == '__main__':
    main()
