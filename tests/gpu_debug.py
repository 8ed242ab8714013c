"""Developer diagnostic (not a test): runs each kernel case in its own subprocess (a trapped kernel kills the
CUDA context) and prints error metrics instead of asserting.  python tests/gpu_debug.py [kernels|unet|all]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONV_SHAPES = [(2, 16, 16, 64, 0, 64), (2, 32, 32, 32, 0, 32), (1, 16, 16, 32, 32, 64), (2, 8, 8, 128, 0, 256),
               (3, 4, 4, 64, 0, 512), (1, 24, 40, 64, 0, 32), (2, 16, 16, 64, 64, 128), (1, 128, 128, 32, 0, 32)]


def case_conv(kind, shape):
    import torch
    from tests import gpu_util as U
    B, H, W, C0, C1, N = shape
    g = torch.Generator(device='cuda').manual_seed(1)
    rb = lambda s: torch.randn(s, generator=g, device='cuda').to(torch.bfloat16)
    x0 = rb((B, H, W, C0))
    x1 = rb((B, H, W, C1)) if C1 else None
    w = torch.randn((3, 3, C0 + C1, N), generator=g, device='cuda') * (2.0 / (9 * (C0 + C1))) ** 0.5
    bias = torch.randn(N, generator=g, device='cuda') * 0.1
    xin = x0 if x1 is None else torch.cat([x0, x1], dim=3)
    bo = 1 if kind.endswith('_bo') else 0
    row = kind.startswith('r')
    kind = kind.replace('_bo', '')
    if row:
        kind = kind[1:]
    conv = (lambda *a, **k: U.conv_row(*a, base_offset_mode=bo, **k)) if row else U.conv_tc
    wg = U.wgrad_row if row else U.wgrad_tc
    if kind == 'fwd':
        out, _, stats = conv(x0, x1, U.pack_fwd(w), bias, N, mode=0, want_stats=True)
        ref = U.ref_conv(xin, w, bias, relu=True)
        s_err = float((stats[:N] - out.double().sum(dim=(0, 1, 2))).abs().max())
        print('  stats max abs err', s_err)
    elif kind == 'dgrad':
        dz = rb((B, H, W, N))
        dx0, dx1, _ = conv(dz, None, U.pack_dgrad(w), None, C0 + C1, mode=2, out_split=C0 if C1 else C0 + C1)
        x = torch.zeros((B, C0 + C1, H, W), device='cuda', requires_grad=True)
        y = torch.nn.functional.conv2d(x, w.to(torch.bfloat16).float().permute(3, 2, 0, 1), padding=1)
        y.backward(dz.float().permute(0, 3, 1, 2))
        ref = x.grad.permute(0, 2, 3, 1)
        out = dx0 if dx1 is None else torch.cat([dx0, dx1], dim=3)
    else:
        dz = rb((B, H, W, N))
        out = wg(x0, x1, dz)
        wv = torch.zeros((N, C0 + C1, 3, 3), device='cuda', requires_grad=True)
        y = torch.nn.functional.conv2d(xin.float().permute(0, 3, 1, 2), wv, padding=1)
        y.backward(dz.float().permute(0, 3, 1, 2))
        ref = wv.grad.permute(2, 3, 1, 0)
    nan = int(torch.isnan(out.float()).sum())
    rel = U.rel_err(torch.nan_to_num(out.float()), ref)
    print('  %s %s: rel %.3e  max %.3e  refmax %.3e  nan %d' % (kind, shape, rel, U.max_err(torch.nan_to_num(out.float()), ref),
                                                                 float(ref.abs().max()), nan))
    if rel > 1e-2:
        o, r = torch.nan_to_num(out.float()).double(), ref.double()
        if kind != 'wgrad':
            flat_o, flat_r = o.reshape(-1, o.shape[-1]), r.reshape(-1, r.shape[-1])
            e = (flat_o - flat_r).abs()
            print('   err by channel block of 8:', [round(float(v), 3) for v in e.reshape(e.shape[0], -1, 8).mean(dim=(0, 2))[:16]])
            print('   err by pixel row %% 8     :', [round(float(e[i::8].mean()), 3) for i in range(8)])
            print('   err by x position (first image row):', [round(float(v), 2) for v in (o - r).abs().mean(dim=3)[0, 0, :16]])
            print('   err by y position (first col):', [round(float(v), 2) for v in (o - r).abs().mean(dim=3)[0, :16, 0]])
            print('   sample got', [round(float(v), 3) for v in flat_o[5, :8]], 'ref', [round(float(v), 3) for v in flat_r[5, :8]])
        else:
            e = (o - r).abs()
            print('   err by tap:', [round(float(v), 3) for v in e.mean(dim=(2, 3)).reshape(-1)])
            print('   err by ci block of 8:', [round(float(v), 3) for v in e.mean(dim=(0, 1, 3)).reshape(-1, 8).mean(dim=1)[:16]])
            print('   err by co block of 8:', [round(float(v), 3) for v in e.mean(dim=(0, 1, 2)).reshape(-1, 8).mean(dim=1)[:16]])
            print('   sample got', [round(float(v), 3) for v in o[1, 1, 3, :8]], 'ref', [round(float(v), 3) for v in r[1, 1, 3, :8]])


def case_unet(precision, dim, depth, batch):
    import numpy as np
    import torch
    from tests.test_gpu_unet import _setup
    from oracle import unet_ref as R
    model, cfg, ws, x, y = _setup(precision, dim, depth, batch, randomize_bn=False)
    heat = model.predict(x, batch_size=batch)
    ref, acts = R.predict(cfg, ws, x, return_acts=True)
    print('  predict %s %d d%d B%d: heat max abs err %.3e mean %.3e' % (precision, dim, depth, batch,
                                                                       np.abs(heat - ref).max(), np.abs(heat - ref).mean()))
    for s in R.layer_specs(cfg)[:-1]:
        for which, key in ((0, '/a'), (1, '/y')):
            t = model.debug_buffer(s.name, which, batch, False)
            if t is None:
                continue
            r = acts[s.name + key]
            g = t.float().cpu().numpy().reshape(r.shape)
            print('    %-16s %s rel %.3e' % (s.name, key, np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)))
    out = R.train_grads(cfg, ws, x, y, return_acts=True)
    loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), apply_optimizer=False).item())
    print('  train: loss %.8f oracle %.8f' % (loss, out['loss']))
    g = model.grads.cpu().numpy()
    for (name, is_state, off, shape), rg in zip(model.tensors, out['grads']):
        if is_state:
            continue
        n = int(np.prod(shape))
        mine = g[off:off + n].reshape(shape).astype(np.float64)
        rg = rg.astype(np.float64)
        cos = (mine * rg).sum() / (np.linalg.norm(mine) * np.linalg.norm(rg) + 1e-300)
        print('    grad %-28s cos %.6f rl2 %.3e  |ref| %.3e' % (name, cos, np.linalg.norm(mine - rg) / (np.linalg.norm(rg) + 1e-300),
                                                               np.linalg.norm(rg)))
    for s in R.layer_specs(cfg)[:-1]:
        t = model.debug_buffer(s.name, 3, batch, True)
        if t is None:
            continue
    print('  launches so far', model.launch_count())


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if what == 'case':
        kind = sys.argv[2]
        args = eval(sys.argv[3])
        if kind == 'unet':
            case_unet(*args)
        else:
            case_conv(kind, args)
        return
    cases = []
    if what in ('kernels', 'all'):
        for kind in ('fwd', 'dgrad', 'wgrad'):
            for s in CONV_SHAPES:
                cases.append((kind, s))
    if what in ('row', 'all'):
        ROW_SHAPES = [(1, 8, 128, 32, 0, 32), (2, 8, 256, 64, 0, 32), (1, 8, 128, 32, 32, 32), (1, 8, 128, 64, 0, 64),
                      (1, 4, 128, 128, 0, 64), (1, 6, 128, 32, 0, 32), (2, 16, 128, 64, 64, 64)]
        for kind in ('rfwd', 'rdgrad', 'rwgrad'):
            for s_ in ROW_SHAPES:
                cases.append((kind, s_))
    if what in ('unet', 'all'):
        cases += [('unet', ('fp32', 32, 2, 3)), ('unet', ('bf16', 32, 2, 4)), ('unet', ('fp32', 64, 4, 2)),
                  ('unet', ('bf16', 64, 4, 4))]
    for kind, args in cases:
        print('== %s %s' % (kind, args), flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), 'case', kind, repr(args)], cwd=ROOT, timeout=120,
                               capture_output=True, text=True)
            print(r.stdout[-12000:], end="")
            if r.returncode != 0:
                print('  EXIT', r.returncode, r.stderr[-1500:])
        except subprocess.TimeoutExpired:
            print('  TIMEOUT')
        sys.stdout.flush()


if __name__ == '__main__':
    main()
