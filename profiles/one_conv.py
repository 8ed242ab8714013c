"""One launch of the halo conv kernel on a named deep-level shape (workload for ncu source-level captures).
usage: python profiles/one_conv.py <B> <H> <W> <C0> <C1> <Cout>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import gpu_util as U  # noqa: E402

B, H, W, C0, C1, N = [int(v) for v in sys.argv[1:7]]
g = torch.Generator(device='cuda').manual_seed(1)
x0 = torch.randn((B, H, W, C0), generator=g, device='cuda').to(torch.bfloat16)
x1 = torch.randn((B, H, W, C1), generator=g, device='cuda').to(torch.bfloat16) if C1 else None
w = torch.randn((3, 3, C0 + C1, N), generator=g, device='cuda') * 0.02
bias = torch.zeros(N, device='cuda')
for _ in range(2):
    out, _, stats = U.conv_halo(x0, x1, U.pack_fwd(w), bias, N, mode=0, want_stats=True)
torch.cuda.synchronize()
print('ok', float(out.float().abs().mean()))
