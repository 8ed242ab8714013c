"""Inference latency by batch size (predict_model.py:89 feeds batch 1): device time per forward (CUDA events), host time
to ISSUE one forward (launch-bound when larger than the device time), and the same forward replayed from a CUDA graph.
usage: python profiles/infer_latency.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import CONFIG  # noqa: E402
from cmr_landmark_detection_b200 import synth  # noqa: E402
from cmr_landmark_detection_b200.models.Unets import create_unet  # noqa: E402


def main():
    model = create_unet(dict(CONFIG))
    print('batch,device_us_per_forward,host_issue_us_per_forward,graph_us_per_forward,launches')
    for B in (1, 2, 4, 16):
        x, _ = synth.make_batch(B, 256, 256, seed=1)
        xd = torch.from_numpy(x).cuda()
        out = torch.empty((B, 256, 256, 2), dtype=torch.float32, device='cuda')
        for _ in range(5):
            model.predict_device(xd, out)
        torch.cuda.synchronize()
        n = 50
        l0 = model.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            model.predict_device(xd, out)
        e1.record()
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        dev_us = e0.elapsed_time(e1) * 1e3 / n
        launches = (model.launch_count() - l0) // n
        graph_us = float('nan')
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    model.predict_device(xd, out)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                model.predict_device(xd, out)
            for _ in range(5):
                g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            graph_us = e0.elapsed_time(e1) * 1e3 / n
        except Exception as ex:   # capture is an experiment here, not a product path
            print('# graph capture failed: %s' % str(ex)[:200])
        print('%d,%.1f,%.1f,%.1f,%d' % (B, dev_us, t_issue * 1e6 / n, graph_us, launches))


if __name__ == '__main__':
    main()
