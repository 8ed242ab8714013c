"""End-to-end parity of the device U-Net path (through create_unet -> C ABI) against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BASE = {'DEPTH': 2, 'FILTERS': 32, 'IMG_CHANNELS': 1, 'MASK_CLASSES': 2, 'BATCH_NORMALISATION': True,
        'BN_FIRST': False, 'ACTIVATION': 'relu', 'PAD': 'same', 'DROPOUT_MIN': 0.0, 'DROPOUT_MAX': 0.0,
        'LEARNING_RATE': 1e-3, 'M_POOL': [2, 2], 'F_SIZE': [3, 3], 'SEED': 7}

# tolerances: heat max-abs / loss rel / gradient cosine + rel-L2 / Adam update rel
TOL = {'fp32': dict(heat=1e-4, loss=1e-5, cos=0.99999, rl2=3e-3, upd=1e-3),
       'bf16': dict(heat=2e-2, loss=1e-2, cos=0.999, rl2=3e-2, upd=5e-2)}


def _setup(precision, dim, depth, batch, randomize_bn=True, seed=0, extra=None):
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.Unets import create_unet
    from oracle import unet_ref as R
    config = dict(BASE, DIM=[dim, dim], DEPTH=depth, PRECISION=precision, **(extra or {}))
    model = create_unet(config)
    cfg = R.cfg_from_config(config)
    ws = R.init_weights(cfg, seed=11 + seed, randomize_bn=randomize_bn)
    model.set_weights(ws)
    x, y = synth.make_batch(batch, dim, dim, seed=5 + seed)
    return model, cfg, ws, x, y


@pytest.mark.parametrize('precision,dim,depth,batch', [('fp32', 32, 2, 3), ('bf16', 32, 2, 3), ('fp32', 64, 4, 2),
                                                       ('bf16', 64, 4, 4), ('bf16', 256, 4, 2), ('bf16', 224, 4, 1)])
def test_predict_matches_oracle(precision, dim, depth, batch):
    from oracle import unet_ref as R
    model, cfg, ws, x, y = _setup(precision, dim, depth, batch)
    heat = model.predict(x, batch_size=batch)
    ref = R.predict(cfg, ws, x)
    assert heat.shape == ref.shape and heat.dtype == np.float32
    t = TOL[precision]
    assert np.abs(heat - ref).max() <= t['heat'], np.abs(heat - ref).max()
    if precision == 'bf16':
        assert np.abs(heat - ref).mean() <= 2e-3


@pytest.mark.parametrize('dim,depth,batch', [(64, 2, 3), (128, 3, 2)])
def test_conv2d_transpose_decoder_matches_oracle(dim, depth, batch):
    """USE_UPSAMPLE=False (KerasLayers.py:762-765): the decoder's Conv2DTranspose(3, strides 2, 'same') layers run on the
    phase-decomposed kernels; kernels are (kh, kw, out, in) in get_weights() order like Keras."""
    from oracle import unet_ref as R
    model, cfg, ws, x, y = _setup('bf16', dim, depth, batch, extra={'USE_UPSAMPLE': False})
    assert cfg.use_upsample is False
    shapes = {n: tuple(s) for n, _, _, s in model.tensors}
    f = BASE['FILTERS']
    assert shapes['dec%d.upconv/kernel' % (depth - 1)] == (3, 3, f, 2 * f)
    assert [tuple(w.shape) for w in model.get_weights()] == [tuple(w.shape) for w in ws]
    heat = model.predict(x, batch_size=batch)
    ref = R.predict(cfg, ws, x)
    assert np.abs(heat - ref).max() <= TOL['bf16']['heat'] and np.abs(heat - ref).mean() <= 2e-3
    # training step: loss vs the fp32 oracle, every gradient tensor vs the oracle restated with bf16 storage points
    model, cfg, ws, x, y = _setup('bf16', dim, depth, batch, randomize_bn=False, extra={'USE_UPSAMPLE': False})
    fp = R.train_grads(cfg, ws, x, y)
    cal = R.train_grads(cfg, ws, x, y, storage='bf16')
    loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                         apply_optimizer=False).item())
    assert abs(loss - fp['loss']) <= TOL['bf16']['loss'] * abs(fp['loss']), (loss, fp['loss'])
    g = model.grads.cpu().numpy()
    for (name, is_state, off, shape), rg, cg in zip(model.tensors, fp['grads'], cal['grads']):
        if is_state or np.linalg.norm(rg) < 1e-12:
            continue
        mine = g[off:off + int(np.prod(shape))].reshape(shape).astype(np.float64)
        e_dev = float(np.linalg.norm(mine - rg) / np.linalg.norm(rg))
        e_cal = float(np.linalg.norm(cg.astype(np.float64) - rg) / np.linalg.norm(rg))
        if e_cal <= 0.3:
            assert e_dev <= 2.0 * e_cal + 0.03, (name, e_dev, e_cal)
        else:
            assert e_dev <= 2.5 * e_cal, (name, e_dev, e_cal)


@pytest.mark.parametrize('dim,depth,batch', [(32, 2, 3), (64, 2, 2), (64, 3, 4)])
def test_conv2d_transpose_decoder_fp32_matches_oracle(dim, depth, batch):
    """USE_UPSAMPLE=False in fp32 parity mode: Conv2DTranspose as a CUDA-core convolution over the virtually zero-stuffed
    low-resolution tensor (conv_simt.cu) -- heat maps, loss and EVERY gradient tensor against the fp32 oracle."""
    from oracle import unet_ref as R
    model, cfg, ws, x, y = _setup('fp32', dim, depth, batch, extra={'USE_UPSAMPLE': False})
    assert [tuple(w.shape) for w in model.get_weights()] == [tuple(w.shape) for w in ws]
    heat = model.predict(x, batch_size=batch)
    ref = R.predict(cfg, ws, x)
    assert np.abs(heat - ref).max() <= TOL['fp32']['heat'], np.abs(heat - ref).max()
    model, cfg, ws, x, y = _setup('fp32', dim, depth, batch, randomize_bn=False, extra={'USE_UPSAMPLE': False})
    out = R.train_grads(cfg, ws, x, y)
    loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                         apply_optimizer=False).item())
    assert abs(loss - out['loss']) <= TOL['fp32']['loss'] * abs(out['loss']), (loss, out['loss'])
    g = model.grads.cpu().numpy()
    for (name, is_state, off, shape), rg in zip(model.tensors, out['grads']):
        if is_state or np.linalg.norm(rg) < 1e-12:
            continue
        mine = g[off:off + int(np.prod(shape))].reshape(shape).astype(np.float64)
        rl2 = float(np.linalg.norm(mine - rg) / np.linalg.norm(rg))
        # depth 3 at 64 x 64 normalises over few values at the bottom: same fp32 conditioning bound as
        # test_train_step_matches_oracle uses for depth > 2
        # sums of dy / dz under BatchNorm nearly cancel (biases, beta, gamma): atomics-order noise shows there first
        assert rl2 <= (1e-2 if name.endswith('/kernel') else 3e-2), (name, rl2)


@pytest.mark.parametrize('momentum,nesterov', [(0.0, False), (0.0, True), (0.9, True), (0.9, False)])
def test_sgd_steps_match_oracle(momentum, nesterov):
    """tf.keras.optimizers.SGD on the device (OPTIMIZER='sgd' and the Adam -> SGD switch of the reference's
    OptimizerChanger): two steps from the device's own gradients against the Keras update formula."""
    from cmr_landmark_detection_b200.runtime.model import SGD
    from oracle import unet_ref as R
    model, cfg, ws, x, y = _setup('fp32', 32, 2, 3, randomize_bn=False)
    model.compile(optimizer=SGD(lr=0.05, momentum=momentum, nesterov=nesterov))
    ref_opt = R.SGD(lr=0.05, momentum=momentum, nesterov=nesterov)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    cur = [w.copy() for w in ws]
    for _ in range(2):
        model.train_step_device(xd, yd, apply_optimizer=False)
        g = model.grads.cpu().numpy()
        grads = [None if st else g[off:off + int(np.prod(shape))].reshape(shape).copy()
                 for (name, st, off, shape) in model.tensors]
        model.apply_gradients()
        new = model.get_weights()
        cur = ref_opt.step(cur, grads)
        for (name, st, off, shape), a, b in zip(model.tensors, new, cur):
            if not st:
                assert np.allclose(a, b, rtol=1e-6, atol=1e-7), name
        # BatchNorm moving statistics moved on the device; carry them over so that both sides keep the same state
        cur = [a if st else b for (name, st, off, shape), a, b in zip(model.tensors, new, cur)]


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_first_layer_mappings_agree(precision, monkeypatch):
    """The Cin = 1 first layer has two thread mappings (4 channels x 8 pixels, and 8 channels x 4 pixels for widths that
    are not multiples of 8): same forward output and the same weight gradient up to summation order."""
    from oracle import unet_ref as R
    out = {}
    for quad in (False, True):
        if quad:
            monkeypatch.setenv('RVIP_C1_QUAD', '1')
        else:
            monkeypatch.delenv('RVIP_C1_QUAD', raising=False)
        model, cfg, ws, x, y = _setup(precision, 32, 2, 3, randomize_bn=False)
        ref = R.train_grads(cfg, ws, x, y)
        loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                             apply_optimizer=False).item())
        assert abs(loss - ref['loss']) <= TOL[precision]['loss'] * abs(ref['loss'])
        name, _, off, shape = model.tensors[0]
        assert name == 'enc0.conv_a/kernel'
        n = int(np.prod(shape))
        out[quad] = (model.debug_buffer('enc0.conv_a', 0, 3, True).float().cpu().numpy(),
                     model.grads.cpu().numpy()[off:off + n].copy())
    assert np.array_equal(out[False][0], out[True][0])              # same FMA order per output element
    g0, g1 = out[False][1], out[True][1]
    # two separate steps: in bf16 mode the atomics-order noise of the BatchNorm reductions upstream flips bf16 roundings of
    # dz (run-to-run ~0.5 % on this ill-conditioned first-layer gradient, same mapping or not).  fp32 mode: the 9 x Cout
    # sums are folded with float atomics (shared, then one global add per block) in scheduling order, and this gradient
    # is a cancelling sum (|result| << sum |partials|), so two runs of the SAME mapping already differ by up to a few
    # 1e-4 relative (1 run in ~3 exceeded 1e-4 on the GPU box); the mapping-independent kernel test
    # (test_gpu_kernels.py, vs torch fp32) holds the arithmetic itself to 2e-3 / 1e-4
    lim = 2e-3 if precision == 'fp32' else 3e-2
    rel = np.linalg.norm(g0 - g1) / (np.linalg.norm(g0) + 1e-30)
    print('first-layer gradient, two mappings: rel diff %.3e (%s)' % (rel, precision))
    assert rel <= lim


@pytest.mark.parametrize('precision,dim,depth,batch', [('fp32', 32, 2, 3), ('bf16', 32, 2, 4), ('fp32', 64, 4, 2),
                                                       ('bf16', 64, 4, 4), ('bf16', 128, 4, 2), ('bf16', 112, 3, 2),
                                                       # BASELINE config C2's own shapes (256 x 256, depth 4, 32 filters):
                                                       # two 128-pixel row tiles per image row, the row-merged level-0
                                                       # weight gradient, the 256 x 256 BatchNorm passes
                                                       ('fp32', 256, 4, 4), ('bf16', 256, 4, 4)])
def test_train_step_matches_oracle(precision, dim, depth, batch):
    from oracle import unet_ref as R
    model, cfg, ws, x, y = _setup(precision, dim, depth, batch, randomize_bn=False)
    t = TOL[precision]
    ref = R.train_grads(cfg, ws, x, y)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    loss = float(model.train_step_device(xd, yd, apply_optimizer=False).item())
    assert abs(loss - ref['loss']) <= t['loss'] * abs(ref['loss']), (loss, ref['loss'])
    g = model.grads.cpu().numpy()
    # Gradients of this BatchNorm net at random init are ill-conditioned: rounding ANY stored tensor to bf16
    # (even only the weights) moves deep-layer gradients by tens of percent in the fp32 oracle itself
    # (DESIGN.md "bf16 gradient conditioning"). So bf16 mode is held to (a) the fp32 oracle on loss, heat maps
    # and the last block's gradients, and (b) the oracle restated with the same bf16 storage points on every
    # gradient tensor. fp32 mode is held to the fp32 oracle everywhere.
    cal = R.train_grads(cfg, ws, x, y, storage='bf16', phased_up=True) if precision == 'bf16' else ref
    # float64 run of the oracle: the yardstick for how much fp32 rounding alone moves each gradient tensor
    ref64 = R.train_grads(cfg, ws, x, y, dtype=torch.float64) if precision == 'fp32' and depth > 2 else None
    last = ('head/', 'dec%d.conv_b/' % (depth - 1))
    for (name, is_state, off, shape), rg, cg in zip(model.tensors, ref['grads'], cal['grads']):
        if is_state:
            continue
        n = int(np.prod(shape))
        mine = g[off:off + n].reshape(shape).astype(np.float64)

        def cmp(r):
            r = r.astype(np.float64)
            nr = np.linalg.norm(r)
            return (float((mine * r).sum() / (np.linalg.norm(mine) * nr + 1e-300)), float(np.linalg.norm(mine - r) / (nr + 1e-300)))
        if np.linalg.norm(rg) < 1e-12:
            continue
        if precision == 'fp32':
            cos, rl2 = cmp(rg)
            # depth-4 nets at 64x64 normalise over as few as 32 values at the bottleneck: fp32 summation-order
            # noise is amplified ~1e3x there (same effect in the oracle run twice with permuted sums)
            lim_cos, lim_rl2 = (t['cos'], t['rl2']) if depth <= 2 else (0.9999, 2e-2)
            if name.endswith('/bias'):      # sum of dz under BatchNorm nearly cancels: atomics-order noise shows
                lim_cos, lim_rl2 = min(lim_cos, 0.9999), max(lim_rl2, 1e-2)
            if ref64 is not None and not (cos >= lim_cos and rl2 <= lim_rl2):
                # ill-conditioned tensor (32 values per channel at the bottleneck of this toy shape): the device may
                # be as far from the float64 truth as 20x the fp32 CPU oracle is (atomics / different summation
                # order, same conditioning; run-to-run 8..10x was measured), never worse than 5 %
                r64 = ref64['grads'][[n for n, *_ in model.tensors].index(name)]
                e_ref = float(np.linalg.norm(rg.astype(np.float64) - r64) / np.linalg.norm(r64))
                e_dev = cmp(r64)[1]
                assert e_dev <= min(20 * e_ref, 5e-2), (name, cos, rl2, e_dev, e_ref)
                continue
            assert cos >= lim_cos and rl2 <= lim_rl2, (name, cos, rl2)
        else:
            # no worse than what bf16 storage itself does to the fp32 oracle (calibration run `cal`)
            e_dev = cmp(rg)[1]
            e_cal = float(np.linalg.norm(cg.astype(np.float64) - rg) / np.linalg.norm(rg))
            if e_cal <= 0.3:
                # two independent bf16 roundings of the same ill-conditioned gradient sit ~sqrt(2) e_cal apart on
                # average; the 32 x 32 toy net's first-layer kernel has been observed at 1.9 e_cal
                assert e_dev <= 2.0 * e_cal + 0.03, ('bf16 path vs calibration', name, e_dev, e_cal)
            else:
                # calibration says this tensor's gradient is rounding-noise dominated (two bf16 runs differ from fp32
                # by e_cal each and from one another by ~sqrt(2) e_cal): require the right magnitude only
                ratio = np.linalg.norm(mine) / np.linalg.norm(rg)
                assert e_dev <= 2.5 * e_cal and 0.3 <= ratio <= 3.0, ('bf16 path vs calibration', name, e_dev, e_cal, ratio)
            if name.startswith(last):
                cos, rl2 = cmp(rg)
                # conv biases under BatchNorm: the gradient is a sum of dz that nearly cancels, so every rounding
                # upstream (here also the single bf16 rounding of the pre-summed up-convolution taps) shows first there
                lim = 0.995 if name.endswith('/bias') else 0.998
                assert cos >= lim and rl2 <= (0.1 if name.endswith('/bias') else 7e-2), ('vs fp32 oracle', name, cos, rl2)
    # BN moving statistics after one step
    new = R.apply_new_stats(cfg, ws, ref['new_stats'])
    mine = model.get_weights()
    for (name, is_state, off, shape), a, b in zip(model.tensors, mine, new):
        if is_state:
            if precision == 'bf16':
                assert np.allclose(a, b, rtol=5e-2, atol=1e-3), name
            else:
                assert np.allclose(a, b, rtol=1e-4, atol=1e-6), name
    # Adam step
    opt = R.Adam(lr=1e-3)
    stepped = opt.step(ws, ref['grads'])
    model.apply_gradients()
    after = model.get_weights()
    if precision == 'fp32' and depth <= 2:
        for (name, is_state, off, shape), a, b, w0, rg in zip(model.tensors, after, stepped, ws, ref['grads']):
            if is_state:
                continue
            # first Adam step: upd = lr * g / (|g| + eps'), eps' = eps / sqrt(1 - beta2) = 3.16e-6, so an element's
            # update moves by lr * eps' * dg / (|g| + eps')^2 when its gradient moves by dg.  The norm-wise checks
            # above hold the tensor to rel-L2 1e-4; single elements carry fp32 summation-order noise of about
            # 1e-6 * sum|terms| ~ 1e-4 * max|g| (atomics), which is what dg allows for.
            eps_h = 1e-7 / np.sqrt(1 - 0.999)
            gabs = np.abs(rg.astype(np.float64))
            big = gabs > 1e-3 * gabs.max()
            dg = 2e-2 * gabs + 2e-4 * gabs.max()
            tol = 5e-3 * np.abs(b - w0) + 1e-3 * eps_h * dg / (gabs + eps_h) ** 2 + 1e-7
            viol = (np.abs((a - w0) - (b - w0)) / tol)[big]
            k = int(np.argmax(viol))
            assert viol[k] <= 1.0, (name, float(viol[k]), float(gabs[big][k]), float(gabs.max()),
                                    float((a - w0)[big][k]), float((b - w0)[big][k]))
    if True:
        # Adam on the device gradients themselves (the optimizer kernel is exact given its input)
        gl = [None if st else g[off:off + int(np.prod(shp))].reshape(shp) for (nm, st, off, shp) in model.tensors]
        mine_step = R.Adam(lr=1e-3).step(ws, gl)
        for (name, is_state, off, shape), a, b in zip(model.tensors, after, mine_step):
            if not is_state:
                assert np.allclose(a, b, rtol=1e-5, atol=1e-7), name


def test_fp32_argmax_bit_exact_and_landmarks_within_half_pixel():
    """north_star: integer argmax bit-exact in fp32 mode (when the oracle's top-2 differ by > 1e-3) and
    landmark coordinates within 0.5 px in both modes."""
    from cmr_landmark_detection_b200.extract import extract_device
    from oracle import extract_ref as ex
    from oracle import unet_ref as R
    for precision in ('fp32', 'bf16'):
        model, cfg, ws, x, y = _setup(precision, 64, 4, 4, seed=3)
        # bias the head so both channels cross the 0.5 threshold somewhere
        heat = model.predict(x, batch_size=4)
        ref = R.predict(cfg, ws, x)
        thr = float(np.quantile(ref, 0.97))
        r = extract_device(torch.from_numpy(heat).cuda(), thr)
        cnt, sr, sc, am, vm = ex.extract_stats(ref, thr)
        yx = r['yx'].cpu().numpy()
        for z in range(ref.shape[0]):
            for c in range(2):
                flat = np.sort(ref[z, :, :, c].ravel())
                if precision == 'fp32' and flat[-1] - flat[-2] > 1e-3:
                    assert int(r['argmax'][z, c]) == int(am[z, c])
                # centroid parity only where the thresholded set is not sitting on the decision boundary
                margin = np.abs(ref[z, :, :, c] - thr).min()
                if cnt[z, c] > 0 and margin > (1e-4 if precision == 'fp32' else 2e-2):
                    assert abs(yx[z, c, 0] - sr[z, c] / cnt[z, c]) <= 0.5
                    assert abs(yx[z, c, 1] - sc[z, c] / cnt[z, c]) <= 0.5


def test_dropout_masks_replayed_in_oracle():
    """Dropout on: export the Philox keep-masks the kernels used and feed them to the oracle."""
    import ctypes as C
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.Unets import create_unet
    from cmr_landmark_detection_b200.runtime import ffi
    from oracle import unet_ref as R
    config = dict(BASE, DIM=[32, 32], DEPTH=2, PRECISION='fp32', DROPOUT_MIN=0.3, DROPOUT_MAX=0.5)
    model = create_unet(config)
    cfg = R.cfg_from_config(config)
    ws = R.init_weights(cfg, seed=21)
    model.set_weights(ws)
    x, y = synth.make_batch(4, 32, 32, seed=9)
    loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                         apply_optimizer=False).item())
    seed = (model._seed * 1000003 + model._step) & (2 ** 64 - 1)
    b = model._bindings[(4, True)]
    masks = {}
    shapes = {'enc0.conv_a': (4, 32, 32, 32), 'enc1.conv_a': (4, 16, 16, 64), 'mid.conv_a': (4, 8, 8, 128),
              'dec0.conv_a': (4, 16, 16, 64), 'dec1.conv_a': (4, 32, 32, 32)}
    names = {'enc0.conv_a': 'enc0', 'enc1.conv_a': 'enc1', 'mid.conv_a': 'mid', 'dec0.conv_a': 'dec0',
             'dec1.conv_a': 'dec1'}
    for lname, shp in shapes.items():
        site, rate = C.c_uint32(), C.c_float()
        ffi.check(ffi.lib().rvip_dropout_site(b.h, lname.encode(), C.byref(site), C.byref(rate)))
        n = int(np.prod(shp))
        keep = torch.empty(n, dtype=torch.uint8, device='cuda')
        ffi.check(ffi.lib().rvip_dropout_mask(C.c_uint64(seed), site.value, rate.value, n, ffi.ptr(keep), None))
        torch.cuda.synchronize()
        masks[names[lname]] = keep.cpu().numpy().reshape(shp)
        assert abs(masks[names[lname]].mean() - (1 - rate.value)) < 0.02
    ref = R.train_grads(cfg, ws, x, y, dropout_masks=masks)
    assert abs(loss - ref['loss']) <= 1e-5 * abs(ref['loss'])
    g = model.grads.cpu().numpy()
    name, is_state, off, shape = model.tensors[0]
    mine = g[off:off + int(np.prod(shape))].reshape(shape)
    # fp32 atomics: the summation order (and with it the 4th digit) changes with the grid shape of the passes
    assert np.linalg.norm(mine - ref['grads'][0]) <= 5e-4 * np.linalg.norm(ref['grads'][0])


def test_weights_roundtrip_and_summary(tmp_path):
    from cmr_landmark_detection_b200.models.Unets import create_unet
    model = create_unet(dict(BASE, DIM=[128, 128], DEPTH=4, PRECISION='bf16'))
    assert model.count_params() == 8641730 and model.n_params == 8635842 and model.n_state == 5888
    ws = model.get_weights()
    assert len(ws) == 118
    p = str(tmp_path / 'model.h5')
    model.save_weights(p)
    assert open(p, 'rb').read(8) == b'\x89HDF\r\n\x1a\n'          # a Keras HDF5 weight file (utils/hdf5_lite.py), not a renamed .npz
    m2 = create_unet(dict(BASE, DIM=[128, 128], DEPTH=4, PRECISION='bf16', SEED=99))
    m2.load_weights(p)
    for a, b in zip(ws, m2.get_weights()):
        assert np.array_equal(a, b)
    # the .npz container (any other extension) still works, and a file of another architecture is refused
    q = str(tmp_path / 'weights')
    model.save_weights(q)
    m3 = create_unet(dict(BASE, DIM=[128, 128], DEPTH=4, PRECISION='bf16', SEED=5))
    m3.load_weights(q)
    for a, b in zip(ws, m3.get_weights()):
        assert np.array_equal(a, b)
    small = create_unet(dict(BASE, DIM=[64, 64], DEPTH=2, PRECISION='bf16'))
    with pytest.raises(ValueError):
        small.load_weights(p)
    lines = []
    model.summary(print_fn=lines.append)
    assert any('8,641,730' in l for l in lines)


def test_fit_reduces_loss_and_callbacks_run():
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.Unets import create_unet
    model = create_unet(dict(BASE, DIM=[32, 32], DEPTH=2, PRECISION='bf16', LEARNING_RATE=2e-3))
    x, y = synth.make_batch(16, 32, 32, seed=4)

    class Seq:
        def __len__(self):
            return 4

        def __getitem__(self, i):
            return x[4 * i:4 * i + 4], y[4 * i:4 * i + 4]

        def on_epoch_end(self):
            pass

    seen = []

    class CB:
        def set_model(self, m):
            self.model = m

        def on_epoch_end(self, epoch, logs):
            seen.append(logs['loss'])
            if epoch == 5:
                self.model.stop_training = True

    h = model.fit(x=Seq(), validation_data=Seq(), epochs=20, callbacks=[CB()], initial_epoch=0, max_queue_size=12,
                  verbose=0)
    assert len(seen) == 6 and h.history['loss'][-1] < h.history['loss'][0]
    assert 'val_loss' in h.history


def test_pipelined_fit_equals_sequential_steps():
    """fit() overlaps batch i+1's staging / H2D with step i: the per-step losses and the final weights must equal
    the same batches pushed one by one through train_on_batch (fp32 mode, dropout off)."""
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.Unets import create_unet
    x, y = synth.make_batch(20, 32, 32, seed=6)
    batches = [(x[4 * i:4 * i + 4], y[4 * i:4 * i + 4]) for i in range(5)]

    def run(pipelined):
        # small learning rate: Adam turns atomics-order noise into sign flips of near-zero gradient elements, which a
        # large step would amplify from step to step (a routing bug, by contrast, changes the losses by > 5 %)
        model = create_unet(dict(BASE, DIM=[32, 32], DEPTH=2, PRECISION='fp32', SEED=3, LEARNING_RATE=1e-4))
        if pipelined:
            losses = model._run_steps(iter(batches))
        else:
            losses = [model.train_on_batch(*b) for b in batches]
        return losses, model.get_weights()
    la, wa = run(True)
    lb, wb = run(False)
    assert len(la) == 5
    # atomics make the gradient sums order-dependent at the 1e-6 level: compare to that, not bitwise
    # the first step sees identical weights; afterwards Adam amplifies the atomics-order noise step by step
    assert abs(la[0] - lb[0]) <= 1e-6 * abs(lb[0]) and np.allclose(la, lb, rtol=5e-3), (la, lb)
    # Adam turns 1e-6 gradient noise on near-zero elements into lr-sized differences: bound mean and max
    for a, b in zip(wa, wb):
        d = np.abs(a.astype(np.float64) - b)
        assert d.mean() <= 2e-4 and d.max() <= 6e-4, (d.mean(), d.max())   # 5 steps x lr 1e-4


def test_c5_topology_matches_oracle():
    """BASELINE config C5 topology (5 levels, 64 base filters: 1024 / 2048 channels at the bottom, 138 M parameters) at
    a size the CPU oracle finishes in seconds: parameter count, inference heat maps and the training loss."""
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.Unets import create_unet
    from oracle import unet_ref as R
    config = dict(BASE, DIM=[64, 64], DEPTH=5, FILTERS=64, PRECISION='bf16')
    model = create_unet(config)
    assert model.count_params() == 138376578
    cfg = R.cfg_from_config(config)
    ws = R.init_weights(cfg, seed=5, randomize_bn=True)
    model.set_weights(ws)
    x, y = synth.make_batch(2, 64, 64, seed=8)
    heat = model.predict(x, batch_size=2)
    ref = R.predict(cfg, ws, x)
    assert np.abs(heat - ref).max() <= TOL['bf16']['heat'], np.abs(heat - ref).max()
    out = R.train_grads(cfg, ws, x, y)
    loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                         apply_optimizer=False).item())
    assert abs(loss - out['loss']) <= TOL['bf16']['loss'] * abs(out['loss']), (loss, out['loss'])


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_bce_dice_loss_matches_oracle(precision):
    """Row N2: the reference's default loss (BceDiceLoss, Loss_and_metrics.py:208-228): loss value, head gradients and a
    deep gradient against the oracle (torch autograd over the restated formula)."""
    from cmr_landmark_detection_b200.models.Loss_and_metrics import BceDiceLoss
    from oracle import unet_ref as R
    model, cfg, ws, x, y = _setup(precision, 32, 2, 4, randomize_bn=False, seed=2)
    model.compile(loss=BceDiceLoss(w_bce=1.0, w_dice=1.0))
    ref = R.train_grads(cfg, ws, x, y, loss_kind='bce_dice', loss_params=dict(w_bce=1.0, w_dice=1.0))
    loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                         apply_optimizer=False).item())
    tol = 1e-5 if precision == 'fp32' else 1e-2
    assert abs(loss - ref['loss']) <= tol * abs(ref['loss']) + 1e-6, (loss, ref['loss'])
    g = model.grads.cpu().numpy()
    lim = 3e-3 if precision == 'fp32' else 7e-2
    checked = 0
    for (name, is_state, off, shape), rg in zip(model.tensors, ref['grads']):
        if is_state or not (name.startswith('head/') or name.startswith('dec1.conv_b/')):
            continue
        if precision == 'bf16' and name == 'dec1.conv_b/bias':
            continue        # sum of dz under BatchNorm nearly cancels: bf16 rounding noise dominates this tensor
        mine = g[off:off + int(np.prod(shape))].reshape(shape).astype(np.float64)
        rl2 = np.linalg.norm(mine - rg) / (np.linalg.norm(rg) + 1e-300)
        assert rl2 <= lim, (name, rl2)
        checked += 1
    assert checked >= 4 if precision == 'bf16' else checked >= 5
    # the config-string route of train_model.py:178
    from cmr_landmark_detection_b200.models.Unets import create_unet
    m2 = create_unet(dict(BASE, DIM=[32, 32], DEPTH=2, PRECISION=precision, LOSS_FUNCTION='BcdDiceLoss'))
    assert m2.loss_kind == 'bce_dice'


def test_fit_with_reference_style_callbacks(tmp_path):
    """train_model.py:95-112 flow: get_callbacks(config) -> model.fit(..., callbacks=...): the best-only checkpoint is
    written, reloads into a fresh model and reproduces the predictions; ReduceLROnPlateau / LR log see optimizer.lr."""
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.Unets import create_unet
    from cmr_landmark_detection_b200.utils.KerasCallbacks import get_callbacks
    config = dict(BASE, DIM=[32, 32], DEPTH=2, PRECISION='bf16', LEARNING_RATE=2e-3, MODEL_PATH=str(tmp_path / 'model'),
                  TENSORBOARD_PATH=str(tmp_path / 'tb'), MONITOR_FUNCTION='loss', SAVE_MODEL_FUNCTION='loss')
    model = create_unet(config)
    x, y = synth.make_batch(8, 32, 32, seed=12)
    h = model.fit(x, y, batch_size=4, epochs=3, callbacks=get_callbacks(config), verbose=0, shuffle=False)
    assert len(h.history['loss']) == 3 and 'lr' in h.history
    m2 = create_unet(dict(config, SEED=123))
    m2.load_weights(os.path.join(config['MODEL_PATH'], 'model.h5'))
    # the checkpoint holds the weights of the best epoch (loss fell every epoch here -> the last one)
    if h.history['loss'][-1] == min(h.history['loss']):
        assert np.array_equal(m2.predict(x, batch_size=4), model.predict(x, batch_size=4))


def test_first_layer_recompute_variant_matches_default(monkeypatch):
    """RVIP_C1_RECOMPUTE=1 (first-layer activation recomputed from the image instead of stored; opt-in because it measured
    slower) must give the same loss and gradients as the default path."""
    from oracle import unet_ref as R
    res = {}
    for flag in ('0', '1'):
        if flag == '1':
            monkeypatch.setenv('RVIP_C1_RECOMPUTE', '1')
        else:
            monkeypatch.delenv('RVIP_C1_RECOMPUTE', raising=False)
        model, cfg, ws, x, y = _setup('fp32', 32, 2, 3, randomize_bn=False, seed=4)
        loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                             apply_optimizer=False).item())
        res[flag] = (loss, model.grads.cpu().numpy().copy())
    assert abs(res['0'][0] - res['1'][0]) <= 1e-6 * abs(res['0'][0])
    # run-to-run, fp32 atomics under BatchNorm's cancelling sums already move the flat gradient by ~5e-3 rel-L2
    g0, g1 = res['0'][1].astype(np.float64), res['1'][1].astype(np.float64)
    cos = float((g0 * g1).sum() / (np.linalg.norm(g0) * np.linalg.norm(g1)))
    assert cos >= 0.9999, cos


def test_full_size_c2_properties():
    """BASELINE config C2 at its full size (batch 32, 256 x 256, depth 4, 32 filters) -- too large for the CPU oracle
    in a test, so size-independent properties instead:
      (1) bf16 and fp32 device paths agree on the loss (1e-2) and on the heat maps (2e-2);
      (2) fp32 path: the analytic gradient predicts a central finite difference of the loss along the gradient
          direction (directional derivative, 2 %);
      (3) BatchNorm: every block output has batch mean beta and variance gamma^2 (fresh init: 0 and 1)."""
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.Unets import create_unet
    cfg = dict(BASE, DIM=[256, 256], DEPTH=4, FILTERS=32)
    x, y = synth.make_batch(32, 256, 256, seed=21)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    m32 = create_unet(dict(cfg, PRECISION='fp32'))
    ws = m32.get_weights()
    loss32 = float(m32.train_step_device(xd, yd, apply_optimizer=False).item())
    g = m32.grads.clone()
    mb = create_unet(dict(cfg, PRECISION='bf16'))
    mb.set_weights(ws)
    lossb = float(mb.train_step_device(xd, yd, apply_optimizer=False).item())
    assert abs(lossb - loss32) <= 1e-2 * abs(loss32), (lossb, loss32)
    # (3) batch statistics of a BN output (training mode buffers of the bf16 model)
    yb = mb.debug_buffer('enc1.conv_a', 1, 32, True).float()
    assert yb is not None
    ch = yb.reshape(-1, 64)
    assert float(ch.mean(dim=0).abs().max()) < 2e-2 and float((ch.var(dim=0, unbiased=False) - 1).abs().max()) < 5e-2
    # (2) directional derivative, fp32
    # along the normalised gradient itself (a random direction in 8.6 M dimensions projects to ~1e-5 of the loss,
    # below what an fp32 loss can resolve); step sized for a ~2e-3 relative change of the loss
    d = g / g.norm()
    eps = 2e-3 * abs(loss32) / float(g.norm())
    p0 = m32.params.clone()
    vals = []
    for s in (+1, -1):
        m32.params.copy_(p0 + s * eps * d)
        m32._version += 1                      # operand copies must be re-derived
        vals.append(float(m32.train_step_device(xd, yd, apply_optimizer=False).item()))
    m32.params.copy_(p0)
    fd = (vals[0] - vals[1]) / (2 * eps)
    an = float((g * d).sum())
    assert abs(fd - an) <= 2e-2 * abs(an) + 1e-6, (fd, an)
    # (1b) heat maps of the two precisions in inference mode (same weights AND moving statistics)
    mb.set_weights(m32.get_weights())
    hb, h32 = mb.predict(x[:4], batch_size=4), m32.predict(x[:4], batch_size=4)
    assert np.abs(hb - h32).max() <= 2e-2


@pytest.mark.parametrize('dim,depth,batch', [((96, 160), 3, 5), ((48, 80), 2, 3), ((128, 384), 2, 2)])
def test_non_square_odd_batch_matches_oracle(dim, depth, batch):
    """Shapes off the bench's beaten path (non-square images, widths that are not multiples of 128, odd batch sizes):
    every level picks a different kernel family (row / halo / generic per-tap) and partial tiles appear."""
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.Unets import create_unet
    from oracle import unet_ref as R
    config = dict(BASE, DIM=list(dim), DEPTH=depth, PRECISION='bf16')
    model = create_unet(config)
    cfg = R.cfg_from_config(config)
    ws = R.init_weights(cfg, seed=31, randomize_bn=True)
    model.set_weights(ws)
    x, y = synth.make_batch(batch, dim[0], dim[1], seed=13)
    heat = model.predict(x, batch_size=batch)
    ref = R.predict(cfg, ws, x)
    assert heat.shape == ref.shape
    assert np.abs(heat - ref).max() <= TOL['bf16']['heat'], np.abs(heat - ref).max()
    out = R.train_grads(cfg, ws, x, y)
    loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                         apply_optimizer=False).item())
    assert abs(loss - out['loss']) <= TOL['bf16']['loss'] * abs(out['loss']), (loss, out['loss'])
    g = model.grads.cpu().numpy()
    last = 'dec%d.conv_b/' % (depth - 1)
    for (name, is_state, off, shape), rg in zip(model.tensors, out['grads']):
        if is_state or not (name.startswith('head/') or name.startswith(last)) or name.endswith('conv_b/bias'):
            continue
        mine = g[off:off + int(np.prod(shape))].reshape(shape).astype(np.float64)
        assert np.linalg.norm(mine - rg) <= 7e-2 * np.linalg.norm(rg), name


def test_bf16_training_trajectory_tracks_fp32_oracle():
    """The bf16 device path must TRAIN like the fp32 oracle, not only match it at random init: 30 Adam steps on a small
    net (cycling three fixed batches, dropout off) -- the two loss curves stay within a stated band of one another --
    and then the per-tensor gradient check is repeated at the trained weights (BatchNorm gamma / beta away from 1 / 0,
    non-zero biases), against SURVEY 8c's bf16 tolerance where the problem is well conditioned.  The measured numbers go
    to gpurun_out/trajectory_grad_parity.json (DESIGN section 2 quotes them)."""
    import json
    from cmr_landmark_detection_b200 import synth
    from oracle import unet_ref as R
    model, cfg, ws, _, _ = _setup('bf16', 32, 2, 4, randomize_bn=False, seed=6)
    batches = [synth.make_batch(4, 32, 32, seed=40 + i) for i in range(3)]
    opt = R.Adam(lr=1e-3)
    cur = [w.copy() for w in ws]
    ref_curve, dev_curve = [], []
    for i in range(30):
        x, y = batches[i % 3]
        out = R.train_grads(cfg, cur, x, y)
        cur = opt.step(R.apply_new_stats(cfg, cur, out['new_stats']), out['grads'])
        ref_curve.append(out['loss'])
        dev_curve.append(model.train_on_batch(x, y))
    ref_curve, dev_curve = np.array(ref_curve), np.array(dev_curve)
    assert ref_curve[-3:].mean() < 0.7 * ref_curve[:3].mean()            # the oracle actually trained
    band = np.abs(dev_curve - ref_curve) / ref_curve
    assert band.max() <= 5e-2, (band.max(), dev_curve.tolist(), ref_curve.tolist())
    # gradient parity at the device's trained weights
    wt = model.get_weights()
    x, y = batches[0]
    ref = R.train_grads(cfg, wt, x, y)
    cal = R.train_grads(cfg, wt, x, y, storage='bf16', phased_up=True)
    model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), apply_optimizer=False)
    g = model.grads.cpu().numpy()
    rows = {}
    for (name, is_state, off, shape), rg, cg in zip(model.tensors, ref['grads'], cal['grads']):
        if is_state or np.linalg.norm(rg) < 1e-12:
            continue
        mine = g[off:off + int(np.prod(shape))].reshape(shape).astype(np.float64)
        rg = rg.astype(np.float64)
        cos = float((mine * rg).sum() / (np.linalg.norm(mine) * np.linalg.norm(rg)))
        rl2 = float(np.linalg.norm(mine - rg) / np.linalg.norm(rg))
        e_cal = float(np.linalg.norm(cg.astype(np.float64) - rg) / np.linalg.norm(rg))
        rows[name] = dict(cos=cos, rel_l2=rl2, cal_rel_l2=e_cal)
    os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out'), exist_ok=True)
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out',
                           'trajectory_grad_parity.json'), 'w') as f:
        json.dump(dict(loss_band_max=float(band.max()), ref_curve=ref_curve.tolist(), dev_curve=dev_curve.tolist(),
                       grads=rows), f, indent=1)
    # SURVEY 8c bf16 tolerance (cos >= 0.999, rel-L2 <= 3e-2) wherever bf16 STORAGE itself allows it (calibration run of
    # the oracle with the same storage points within 1e-2 of its own fp32 self); elsewhere no worse than 2x calibration.
    # (The trained weights differ a little from run to run -- atomics order over 30 bf16 steps -- so a tensor whose
    # calibration error sits at the class boundary changes class between runs: with the boundary at 1.5e-2 one run in
    # ~10 put a tensor with rel-L2 3.1e-2 into the strict class.  At 1e-2 the strict class is the head and the last
    # block, whose errors are 10x inside the tolerance.)
    n_strict = 0
    for name, v in rows.items():
        if v['cal_rel_l2'] <= 1e-2:
            assert v['cos'] >= 0.999 and v['rel_l2'] <= 3e-2, (name, v)
            n_strict += 1
        else:
            assert v['rel_l2'] <= 2.0 * v['cal_rel_l2'] + 0.03, (name, v)
    assert n_strict >= 1, rows


@pytest.mark.parametrize('dim,depth', [(128, 2), (128, 3)])
def test_fused_bn_backward_sums_variant_matches_default(dim, depth, monkeypatch):
    """RVIP_BNRED_FUSION=all (opt-in: the dgrad epilogues also produce the BatchNorm-backward sums of the block they feed,
    replaying the dropout mask -- measured slower) must give the same gradients as the default path with
    its separate statistics pass.  Dropout ON, same seed, bf16: both row (128 px wide) and halo kernels take part."""
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.Unets import create_unet
    res = {}
    for flag in (None, 'all'):
        if flag:
            monkeypatch.setenv('RVIP_BNRED_FUSION', flag)
        else:
            monkeypatch.delenv('RVIP_BNRED_FUSION', raising=False)
        model = create_unet(dict(BASE, DIM=[dim, dim], DEPTH=depth, PRECISION='bf16', DROPOUT_MIN=0.3, DROPOUT_MAX=0.5))
        x, y = synth.make_batch(4, dim, dim, seed=14)
        loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                             apply_optimizer=False).item())
        res[flag] = (loss, model.grads.cpu().numpy().astype(np.float64), model.tensors)
    assert res[None][0] == res['all'][0]            # the forward pass is untouched
    g0, g1, tensors = res[None][1], res['all'][1], res[None][2]
    worst = 1.0
    for name, is_state, off, shape in tensors:
        if is_state:
            continue
        n = int(np.prod(shape))
        a, b = g0[off:off + n], g1[off:off + n]
        if np.linalg.norm(a) < 1e-12:
            continue
        # the fused sums see the fp32 accumulator where the separate pass sees its bf16 rounding, and the atomics order
        # differs: BatchNorm's cancelling sums turn that into a few per cent on the deepest-path tensors
        cos = float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b)))
        worst = min(worst, cos)
        if name.startswith(('head/', 'dec%d.' % (depth - 1))):
            assert cos >= 0.999, (name, cos)
    # run-to-run, two bf16 steps of the DEFAULT path already differ by a few per cent on the deepest-path tensors (atomics
    # order under BatchNorm's cancelling sums, test_first_layer_mappings_agree); a wrong mask or a missed tile would be O(1)
    assert worst >= 0.97, worst


def test_deferred_weight_gradients_variant_matches_default(monkeypatch):
    """RVIP_DEFER_WGRAD (opt-in schedule: deep-level weight gradients keep dz in buffers of their own and are queued when
    backward reaches the wide encoder levels; measured neutral) changes WHEN kernels run, never what they
    compute: same loss, same gradients up to the run-to-run atomics-order noise, and the optimizer step still sees every
    bucket complete (two Adam steps give the same weights)."""
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.Unets import create_unet
    res = {}
    for flag in (None, 'dec0.conv_a,dec1.conv_a,mid.conv_b,enc3.conv_b,enc2.conv_a'):
        if flag:
            monkeypatch.setenv('RVIP_DEFER_WGRAD', flag)
        else:
            monkeypatch.delenv('RVIP_DEFER_WGRAD', raising=False)
        model = create_unet(dict(BASE, DIM=[64, 64], DEPTH=4, PRECISION='bf16'))
        x, y = synth.make_batch(4, 64, 64, seed=21)
        xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
        loss = float(model.train_step_device(xd, yd, apply_optimizer=False).item())
        g = model.grads.cpu().numpy().astype(np.float64)
        for _ in range(2):
            model.train_step_device(xd, yd)
        res[flag] = (loss, g, model.params.cpu().numpy().astype(np.float64), model.tensors)
    (l0, g0, p0, tensors), (l1, g1, p1, _) = res[None], res[[k for k in res if k][0]]
    assert abs(l0 - l1) <= 2e-3 * abs(l0), (l0, l1)      # the forward pass is untouched; its statistics atomics reorder
    for name, is_state, off, shape in tensors:
        if is_state:
            continue
        n = int(np.prod(shape))
        a, b = g0[off:off + n], g1[off:off + n]
        if np.linalg.norm(a) < 1e-12:
            continue
        cos = float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b)))
        assert cos >= 0.97, (name, cos)                      # same bound as two runs of the default path
        if name.endswith('/kernel'):
            # a weight gradient that was never queued (or an optimizer that ran before it) would leave the kernel unchanged
            # or far off: Adam moves every element by ~lr per step
            wa, wb = p0[off:off + n], p1[off:off + n]
            assert np.abs(wa - wb).max() <= 4.1e-3, (name, float(np.abs(wa - wb).max()))
