"""Structural pins + self-consistency of the U-Net oracle (the numeric half is 'parity unpinned':
TensorFlow is not importable; see oracle/unet_ref.py header)."""
import json
import os

import numpy as np
import torch

from oracle import unet_ref as R

CFG = {'DIM': [128, 128], 'DEPTH': 4, 'FILTERS': 32, 'IMG_CHANNELS': 1, 'MASK_CLASSES': 2,
       'BATCH_NORMALISATION': True, 'BN_FIRST': False, 'DROPOUT_MIN': 0.3, 'DROPOUT_MAX': 0.5}


def test_param_counts_match_reference_summary(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, 'unet_summary.json')))
    cfg = R.cfg_from_config(CFG)
    total, train, non = R.count_params(cfg)
    assert total == gold['totals']['Total params'] == 8641730
    assert train == gold['totals']['Trainable params'] == 8635842
    assert non == gold['totals']['Non-trainable params'] == 5888
    # per-layer counts in creation order
    ref = [(l['type'], l['params']) for l in gold['layers'] if l['params'] > 0]
    mine = []
    for s in R.layer_specs(cfg):
        mine.append(('Conv2D', s.k * s.k * s.cin * s.cout + s.cout))
        if s.bn:
            mine.append(('BatchNormalization', 4 * s.cout))
    assert len(ref) == len(mine) == 23 + 18
    for (rt, rp), (mt, mp) in zip(ref, mine):
        assert mt.startswith(rt[:6]) and rp == mp


def test_dropout_schedules():
    assert R.cfg_from_config(CFG).dropouts == [0.3, 0.4, 0.4, 0.5]
    c5 = dict(CFG, DEPTH=5)
    assert R.cfg_from_config(c5).dropouts == [0.3, 0.4, 0.4, 0.4, 0.5]
    d5 = R.cfg_from_config(dict(CFG, DEPTH=5, FILTERS=64, DIM=[512, 512]))
    assert R.count_params(d5)[0] == 138376578 and R.count_params(d5)[2] == 24064


def test_use_upsample_string_default_is_truthy():
    assert R.cfg_from_config(CFG).use_upsample is True
    assert R.cfg_from_config(dict(CFG, USE_UPSAMPLE=False)).use_upsample is False


def _small():
    cfg = R.cfg_from_config(dict(CFG, DIM=[16, 16], DEPTH=2, FILTERS=4, DROPOUT_MIN=0.0, DROPOUT_MAX=0.0))
    ws = R.init_weights(cfg, seed=3, randomize_bn=True)
    rng = np.random.default_rng(0)
    x = rng.random((3, 16, 16, 1)).astype(np.float32)
    t = rng.random((3, 16, 16, 2)).astype(np.float32)
    return cfg, ws, x, t


def test_numpy_restatement_of_block_ops():
    """conv 'same' / ReLU->BN / maxpool / nearest-up / concat order restated in plain numpy."""
    cfg, ws, x, t = _small()
    heat, acts = R.predict(cfg, ws, x, dtype=torch.float64, return_acts=True)
    k, b, g, be, mm, mv = [w.astype(np.float64) for w in ws[:6]]
    xp = np.pad(x.astype(np.float64), ((0, 0), (1, 1), (1, 1), (0, 0)))
    z = np.zeros((3, 16, 16, 4))
    for dy in range(3):
        for dx in range(3):
            z += xp[:, dy:dy + 16, dx:dx + 16, :] @ k[dy, dx]
    a = np.maximum(z + b, 0)
    y = g * (a - mm) / np.sqrt(mv + 1e-3) + be
    assert np.abs(acts['enc0.conv_a/y'] - y).max() < 1e-12
    assert heat.shape == (3, 16, 16, 2) and (heat > 0).all() and (heat < 1).all()


def test_finite_difference_gradients_float64():
    cfg, ws, x, t = _small()
    out = R.train_grads(cfg, ws, x, t, dtype=torch.float64)
    rng = np.random.default_rng(1)
    tm = R.trainable_mask(cfg)
    checked = 0
    errs = []
    for i, w in enumerate(ws):
        if not tm[i]:
            continue
        for _ in range(2):
            idx = tuple(rng.integers(0, s) for s in w.shape)
            eps = 1e-5
            wp = [v.astype(np.float64) for v in ws]
            wm = [v.astype(np.float64) for v in ws]
            wp[i][idx] += eps
            wm[i][idx] -= eps
            lp = R.train_grads(cfg, wp, x, t, dtype=torch.float64)['loss']
            lm = R.train_grads(cfg, wm, x, t, dtype=torch.float64)['loss']
            fd = (lp - lm) / (2 * eps)
            g = out['grads'][i][idx]
            errs.append(abs(fd - g) / max(abs(g), 1e-3))
            checked += 1
    assert checked > 20
    # ReLU / max-pool kinks inside the +-eps interval perturb a few probes; the bulk must be exact
    errs = np.sort(np.array(errs))
    assert np.median(errs) < 1e-6 and errs[int(0.8 * len(errs))] < 1e-4 and errs[-1] < 5e-2, errs


def test_manual_block_backward_matches_autograd():
    rng = np.random.default_rng(5)
    a_pre = rng.standard_normal((200, 6))
    gamma = rng.uniform(0.5, 1.5, 6)
    dy = rng.standard_normal((200, 6))
    z = torch.tensor(a_pre, requires_grad=True)
    a = torch.relu(z)
    mu, var = a.mean(0), a.var(0, unbiased=False)
    y = torch.tensor(gamma) * (a - mu) / torch.sqrt(var + R.BN_EPS)
    (y * torch.tensor(dy)).sum().backward()
    dz, dg, db = R.manual_block_backward(np.maximum(a_pre, 0), dy, gamma)
    assert np.abs(dz - z.grad.numpy()).max() < 1e-12


def test_maxpool_gradient_goes_to_first_max():
    """Appendix C.5: ties -> first maximal element in row-major window order."""
    x = torch.zeros(1, 1, 2, 2, requires_grad=True)
    torch.nn.functional.max_pool2d(x, 2, 2).sum().backward()
    assert x.grad.flatten().tolist() == [1.0, 0.0, 0.0, 0.0]


def test_adam_keras_form_and_dp_equivalence():
    cfg, ws, x, t = _small()
    x4 = np.concatenate([x, x[:1]], 0)
    t4 = np.concatenate([t, t[:1]], 0)
    dp = R.data_parallel_grads(cfg, ws, x4, t4, world=2)
    # summed-over-replicas gradient of (local sum / global batch) == mean of local-mean grads
    o0 = R.train_grads(cfg, ws, x4[:2], t4[:2])
    o1 = R.train_grads(cfg, ws, x4[2:], t4[2:])
    for g, a, b in zip(dp['grads'], o0['grads'], o1['grads']):
        if g is not None:
            assert np.allclose(g, 0.5 * (a + b), atol=1e-7)
    opt = R.Adam(lr=1e-3)
    new = opt.step(ws, dp['grads'])
    i = 0
    g = dp['grads'][0]
    # first Adam step moves every coordinate by ~lr * sign(g)
    step = new[0] - ws[0]
    # epsilon-hat form, t=1: step = -lr * g / (|g| + eps / sqrt(1 - b2))
    expect = -1e-3 * g / (np.abs(g) + 1e-7 / np.sqrt(1 - 0.999))
    assert np.allclose(step, expect, rtol=1e-4, atol=1e-9)


def test_inplane_weights_match_reference_loop():
    w = R.inplane_weights(10, 10)
    assert w[0].max() == 0 and w[:, 0].max() == 0
    assert w[4, 4] == 100.0 and abs(w[1, 1] - 25.0) < 1e-6 and abs(w[2, 5] - 50.0) < 1e-6


def test_bce_dice_oracle_formula():
    """loss_torch('bce_dice') against a direct numpy evaluation of Loss_and_metrics.py:165-171, 208-228."""
    import torch
    from oracle import unet_ref as R
    rng = np.random.default_rng(3)
    p = rng.uniform(0.001, 0.999, size=(2, 2, 6, 5))
    p[0, 0, 0, 0], p[0, 1, 1, 1] = 0.0, 1.0          # exercise the clipping
    t = (rng.random((2, 2, 6, 5)) < 0.2).astype(np.float64) * rng.uniform(0.5, 1.0, size=(2, 2, 6, 5))
    e = 1e-7
    pc = np.clip(p, e, 1 - e)
    bce = -(t * np.log(pc + e) + (1 - t) * np.log(1 - pc + e)).mean(axis=1)
    dice = (2 * (t * p).sum() + 1) / (t.sum() + p.sum() + 1)
    want = (0.5 * bce - 1.0 * dice).mean()
    got = float(R.loss_torch(torch.tensor(p), torch.tensor(t), 'bce_dice', w_bce=0.5, w_dice=1.0))
    assert abs(got - want) < 1e-12


def test_phased_upconv_restatement_is_the_same_linear_map():
    """UpSampling2D(2) -> Conv3x3 == four 2x2-tap convolutions on the low-resolution tensor with pre-summed taps: the
    calibration restatement of the device's phase-decomposed up-convolution must equal the plain path when nothing is
    rounded (float64, rounding hooks disabled), forward and gradients."""
    import torch
    import torch.nn.functional as F
    from oracle import unet_ref as R

    class Spec:
        name, cin, cout = 'dec0.upconv', 64, 32

    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 16, 32, dtype=torch.float64, generator=g, requires_grad=True)
    k = torch.randn(3, 3, 64, 32, dtype=torch.float64, generator=g, requires_grad=True)
    b = torch.randn(32, dtype=torch.float64, generator=g, requires_grad=True)
    assert R.phased_upconv_eligible(16, 32, 64, 32) and R.phased_upconv_eligible(14, 14, 128, 64)
    assert not R.phased_upconv_eligible(16, 16, 32, 32) and not R.phased_upconv_eligible(16, 16, 64, 48)
    fwd, bwd = R._RoundFwd.apply, R._RoundBwd.apply
    try:
        R._RoundFwd.apply = staticmethod(lambda v: v)
        R._RoundBwd.apply = staticmethod(lambda v: v)
        u = R._upconv_phased_bf16(x, R._Params([k, b]), Spec, {})
    finally:
        R._RoundFwd.apply, R._RoundBwd.apply = fwd, bwd
    ref = torch.relu(R._conv(F.interpolate(x, scale_factor=2, mode='nearest'), k, b))
    assert float((u - ref).abs().max()) < 1e-11
    w = torch.randn(ref.shape, dtype=torch.float64, generator=g)
    g1 = torch.autograd.grad((u * w).sum(), [x, k, b])
    g2 = torch.autograd.grad((ref * w).sum(), [x, k, b])
    for a_, b_ in zip(g1, g2):
        assert float((a_ - b_).abs().max()) < 1e-9


def test_conv2d_transpose_restatement():
    """USE_UPSAMPLE falsy (KerasLayers.py:762-765): Conv2DTranspose(3, strides 2, 'same') [TF-2.3 padding: the whole unit of
    SAME padding sits at the end] is the scatter out[2i + k] += x[i] w[k] cropped to 2h; per output parity it is a
    single-tap filter on the low-resolution tensor -- parity 0: tap 2 on neighbour i - 1 and tap 0 on i, parity 1: tap 1
    on i.  That identity is what the device's phase kernels implement; kernels are (kh, kw, out, in) and the parameter
    count equals the UpSampling2D + Conv2D variant's."""
    import torch
    from oracle import unet_ref as R
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 5, 6, dtype=torch.float64, generator=g)
    k = torch.randn(3, 3, 4, 3, dtype=torch.float64, generator=g)          # (kh, kw, out, in)
    full = R._tconv(x, k, None)
    assert full.shape == (2, 4, 10, 12)
    # direct scatter definition
    ref = torch.zeros(2, 4, 11, 13, dtype=torch.float64)
    for i in range(5):
        for j in range(6):
            for ky in range(3):
                for kx in range(3):
                    ref[:, :, 2 * i + ky, 2 * j + kx] += torch.einsum('oi,ni->no', k[ky, kx], x[:, :, i, j])
    assert torch.allclose(full, ref[:, :, :10, :12], atol=1e-12)
    tap = {(0, 0): 2, (0, 1): 0, (1, 0): 1}          # (parity, neighbour) -> tap index; (1, 1) sees nothing
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))
    for a in range(2):
        for b in range(2):
            acc = torch.zeros(2, 4, 5, 6, dtype=torch.float64)
            for r in range(2):
                for s in range(2):
                    if (a, r) in tap and (b, s) in tap:
                        acc += torch.einsum('oi,nihw->nohw', k[tap[(a, r)], tap[(b, s)]], xp[:, :, a + r:a + r + 5, b + s:b + s + 6])
            assert torch.allclose(acc, full[:, :, a::2, b::2], atol=1e-12)
    up = R.cfg_from_config(dict(CFG, USE_UPSAMPLE=True))
    tr = R.cfg_from_config(dict(CFG, USE_UPSAMPLE=False))
    assert R.count_params(up) == R.count_params(tr)
    shapes = dict(R.weight_shapes(tr))
    spec = [s for s in R.layer_specs(tr) if s.name == 'dec0.upconv'][0]
    assert spec.transposed and shapes['dec0.upconv/kernel'] == (3, 3, spec.cout, spec.cin)
    # float64 finite differences through a small transposed-decoder net
    cfg = R.cfg_from_config({'DIM': [16, 16], 'DEPTH': 2, 'FILTERS': 4, 'IMG_CHANNELS': 1, 'MASK_CLASSES': 2,
                             'BATCH_NORMALISATION': True, 'USE_UPSAMPLE': False, 'DROPOUT_MIN': 0.0, 'DROPOUT_MAX': 0.0})
    ws = R.init_weights(cfg, seed=2, randomize_bn=True)
    rng = np.random.default_rng(4)
    xs, ts = rng.random((2, 16, 16, 1)), rng.random((2, 16, 16, 2))
    out = R.train_grads(cfg, ws, xs, ts, dtype=torch.float64)
    names = [n for n, _ in R.weight_shapes(cfg)]
    i = names.index('dec1.upconv/kernel')
    errs = []
    for _ in range(6):
        idx = tuple(rng.integers(0, s) for s in ws[i].shape)
        wp = [v.astype(np.float64) for v in ws]
        wm = [v.astype(np.float64) for v in ws]
        wp[i][idx] += 1e-5
        wm[i][idx] -= 1e-5
        fd = (R.train_grads(cfg, wp, xs, ts, dtype=torch.float64)['loss'] -
              R.train_grads(cfg, wm, xs, ts, dtype=torch.float64)['loss']) / 2e-5
        errs.append(abs(fd - out['grads'][i][idx]) / max(abs(out['grads'][i][idx]), 1e-4))
    assert np.median(errs) < 1e-5
