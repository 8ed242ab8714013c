"""Config-reachable block variants of conv_layer_fn (src/models/KerasLayers.py:660-693) on the device path against the
fp32 oracle: BN_FIRST=true (Conv -> BN -> ReLU, :681-685) and BATCH_NORMALISATION=false (Conv -> ReLU, :684/:691 skipped),
in fp32 parity mode and on the bf16 tensor-core path, inference and one training step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BASE = {'FILTERS': 32, 'IMG_CHANNELS': 1, 'MASK_CLASSES': 2, 'ACTIVATION': 'relu', 'PAD': 'same', 'DROPOUT_MIN': 0.0,
        'DROPOUT_MAX': 0.0, 'LEARNING_RATE': 1e-3, 'M_POOL': [2, 2], 'F_SIZE': [3, 3], 'SEED': 7}
VARIANTS = {'bn_first': dict(BATCH_NORMALISATION=True, BN_FIRST=True),
            'no_bn': dict(BATCH_NORMALISATION=False, BN_FIRST=False)}


def _setup(variant, precision, dim, depth, batch, randomize_bn, extra=None):
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.Unets import create_unet
    from oracle import unet_ref as R
    config = dict(BASE, DIM=[dim, dim], DEPTH=depth, PRECISION=precision, **VARIANTS[variant], **(extra or {}))
    model = create_unet(config)
    cfg = R.cfg_from_config(config)
    ws = R.init_weights(cfg, seed=23, randomize_bn=randomize_bn)
    if variant == 'no_bn':
        # without BatchNorm the he_normal activations grow with depth and saturate the sigmoid head: damp the kernels a little
        ws = [w * 0.7 if w.ndim == 4 else w for w in ws]
    model.set_weights(ws)
    x, y = synth.make_batch(batch, dim, dim, seed=17)
    return model, cfg, ws, x, y


@pytest.mark.parametrize('variant', ['bn_first', 'no_bn'])
@pytest.mark.parametrize('precision,dim,depth,batch', [('fp32', 32, 2, 3), ('bf16', 64, 3, 4), ('bf16', 128, 2, 2)])
def test_variant_predict_and_train_step_match_oracle(variant, precision, dim, depth, batch):
    from oracle import unet_ref as R
    model, cfg, ws, x, y = _setup(variant, precision, dim, depth, batch, randomize_bn=True)
    assert cfg.bn_first == (variant == 'bn_first') and cfg.batch_norm == (variant != 'no_bn')
    names = [n for n, *_ in model.tensors]
    assert any('/bn/gamma' in n for n in names) == (variant != 'no_bn')
    assert [tuple(w.shape) for w in model.get_weights()] == [tuple(w.shape) for w in ws]
    heat = model.predict(x, batch_size=batch)
    ref = R.predict(cfg, ws, x)
    assert np.abs(heat - ref).max() <= (1e-4 if precision == 'fp32' else 2e-2), np.abs(heat - ref).max()
    # training step (batch statistics)
    model, cfg, ws, x, y = _setup(variant, precision, dim, depth, batch, randomize_bn=False)
    out = R.train_grads(cfg, ws, x, y)
    loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                         apply_optimizer=False).item())
    assert abs(loss - out['loss']) <= (1e-5 if precision == 'fp32' else 1e-2) * abs(out['loss']), (loss, out['loss'])
    g = model.grads.cpu().numpy()
    last = ('head/', 'dec%d.conv_b/' % (depth - 1))
    checked = 0
    for (name, is_state, off, shape), rg in zip(model.tensors, out['grads']):
        if is_state or np.linalg.norm(rg) < 1e-12:
            continue
        mine = g[off:off + int(np.prod(shape))].reshape(shape).astype(np.float64)
        if variant == 'bn_first' and name.endswith('/bias') and not name.startswith(('head/', 'dec0.upconv', 'dec1.upconv',
                                                                                       'dec2.upconv')):
            # z = conv + b feeds BatchNorm directly: d loss / d b = sum dz is EXACTLY zero in exact arithmetic (BatchNorm
            # backward removes the mean); both sides hold rounding noise only
            assert np.linalg.norm(mine) <= 1e-3 * np.abs(g).max() * np.sqrt(mine.size), (name, np.linalg.norm(mine))
            continue
        rl2 = float(np.linalg.norm(mine - rg) / np.linalg.norm(rg))
        cos = float((mine * rg).sum() / (np.linalg.norm(mine) * np.linalg.norm(rg)))
        if precision == 'fp32':
            # conv biases under BatchNorm: a sum of dz that nearly cancels (atomics-order noise shows there first)
            lim = 1e-2 if name.endswith('/bias') and variant != 'no_bn' else 3e-3
            assert rl2 <= lim and cos >= 0.9999, (name, rl2, cos)
            checked += 1
        elif name.startswith(last) and not (name.endswith('conv_b/bias') and variant != 'no_bn'):
            assert rl2 <= 7e-2 and cos >= 0.997, (name, rl2, cos)
            checked += 1
    assert checked >= (10 if precision == 'fp32' else 3), checked
    if variant == 'bn_first':
        new = R.apply_new_stats(cfg, ws, out['new_stats'])
        tol = dict(rtol=1e-4, atol=1e-6) if precision == 'fp32' else dict(rtol=5e-2, atol=2e-3)
        for (name, is_state, off, shape), a, b in zip(model.tensors, model.get_weights(), new):
            if is_state:
                assert np.allclose(a, b, **tol), name


@pytest.mark.parametrize('variant', ['bn_first', 'no_bn'])
def test_variant_with_dropout_replays_masks(variant):
    """Dropout between the convs of a block (KerasLayers.py:718,772) with the variant blocks, fp32: export the keep-masks the
    kernels used and feed them to the oracle."""
    import ctypes as C
    from cmr_landmark_detection_b200.runtime import ffi
    from oracle import unet_ref as R
    model, cfg, ws, x, y = _setup(variant, 'fp32', 32, 2, 4, randomize_bn=False,
                                  extra=dict(DROPOUT_MIN=0.3, DROPOUT_MAX=0.5))
    loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                         apply_optimizer=False).item())
    seed = (model._seed * 1000003 + model._step) & (2 ** 64 - 1)
    b = model._bindings[(4, True)]
    shapes = {'enc0.conv_a': ('enc0', (4, 32, 32, 32)), 'enc1.conv_a': ('enc1', (4, 16, 16, 64)),
              'mid.conv_a': ('mid', (4, 8, 8, 128)), 'dec0.conv_a': ('dec0', (4, 16, 16, 64)),
              'dec1.conv_a': ('dec1', (4, 32, 32, 32))}
    masks = {}
    for lname, (key, shp) in shapes.items():
        site, rate = C.c_uint32(), C.c_float()
        ffi.check(ffi.lib().rvip_dropout_site(b.h, lname.encode(), C.byref(site), C.byref(rate)))
        n = int(np.prod(shp))
        keep = torch.empty(n, dtype=torch.uint8, device='cuda')
        ffi.check(ffi.lib().rvip_dropout_mask(C.c_uint64(seed), site.value, rate.value, n, ffi.ptr(keep), None))
        torch.cuda.synchronize()
        masks[key] = keep.cpu().numpy().reshape(shp)
    ref = R.train_grads(cfg, ws, x, y, dropout_masks=masks)
    assert abs(loss - ref['loss']) <= 1e-5 * abs(ref['loss']), (loss, ref['loss'])
    g = model.grads.cpu().numpy()
    for (name, is_state, off, shape), rg in zip(model.tensors, ref['grads']):
        if is_state or not name.endswith('/kernel'):
            continue
        mine = g[off:off + int(np.prod(shape))].reshape(shape)
        # ReLU / max-pool decisions on values within fp32 rounding of a tie flip between the two implementations: a few
        # 1e-3 on the deepest-path tensors (the same effect bounds test_train_step_matches_oracle at depth > 2)
        assert np.linalg.norm(mine - rg) <= 1e-2 * np.linalg.norm(rg), (name, np.linalg.norm(mine - rg) / np.linalg.norm(rg))
