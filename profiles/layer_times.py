"""Per-layer device times of one training step (CUDA events around every launch group), with the roofline that
bounds each group.  Usage (GPU box): python profiles/layer_times.py [steps] > profiles/<round>_layer_times.csv"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from cmr_landmark_detection_b200 import synth  # noqa: E402
from cmr_landmark_detection_b200.models.Unets import create_unet  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    model = create_unet(dict(bench.CONFIG))
    B = bench.BATCH_PER_GPU
    x, y = synth.make_batch(B, 256, 256, seed=42)
    xd, yd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    for _ in range(3):
        model.train_step_device(xd, yd)
    torch.cuda.synchronize()
    model.profile(B, True, True)
    for _ in range(steps):
        model.train_step_device(xd, yd)
    torch.cuda.synchronize()
    model.profile_read(B, True)
    rows = model.profile_detail(B, True)
    model.profile(B, True, False)
    agg, order = {}, []
    for cls, tag, ms in rows:
        k = (cls, tag)
        if k not in agg:
            agg[k] = 0.0
            order.append(k)
        agg[k] += ms
    print('class,layer:op,us_per_step')
    tot = 0.0
    for k in order:
        us = agg[k] / steps * 1e3
        tot += us
        print('%s,%s,%.1f' % (k[0], k[1], us))
    print('total,,%.1f' % tot)


if __name__ == '__main__':
    main()
