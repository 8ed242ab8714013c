"""Parity of the loss variants, the validation path and the Dice metrics of the device U-Net against the CPU oracle
(src/models/Loss_and_metrics.py:40-89 loss_with_zero_mask, :124-171 dice_coef*; train_model.py:54-59, 105-112)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BASE = {'DEPTH': 2, 'FILTERS': 32, 'IMG_CHANNELS': 1, 'MASK_CLASSES': 2, 'BATCH_NORMALISATION': True,
        'BN_FIRST': False, 'ACTIVATION': 'relu', 'PAD': 'same', 'DROPOUT_MIN': 0.0, 'DROPOUT_MAX': 0.0,
        'LEARNING_RATE': 1e-3, 'M_POOL': [2, 2], 'F_SIZE': [3, 3], 'SEED': 7}


def _setup(precision, dim, depth, batch, seed=0, extra=None, randomize_bn=False):
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


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('kind', ['masked', 'weighted'])
def test_masked_and_weighted_mse_match_oracle(kind, precision):
    """loss_with_zero_mask(mse, mask_smaller_than, weight_inplane): loss value and gradients (head exactly, the last
    block's tensors, and in fp32 every tensor) against oracle.loss_torch + autograd.  The mask threshold is raised so
    that a sizeable share of the pixels is masked out (the synthetic targets are joint-min-max heat maps)."""
    from cmr_landmark_detection_b200.models.Loss_and_metrics import loss_with_zero_mask, mse
    from oracle import unet_ref as R
    model, cfg, ws, x, y = _setup(precision, 32, 2, 4, seed=3)
    thr = 0.05
    model.compile(loss=loss_with_zero_mask(mse, mask_smaller_than=thr, weight_inplane=kind == 'weighted', xy_shape=32))
    assert model.loss_kind == kind
    frac = float((y > thr).any(axis=-1).mean())
    assert 0.02 < frac < 0.9, frac
    whw = R.inplane_weights(32, 32) if kind == 'weighted' else None
    ref = R.train_grads(cfg, ws, x, y, loss_kind=kind, weights_hw=whw, loss_params=dict(mask_smaller_than=thr))
    plain = R.train_grads(cfg, ws, x, y)
    assert abs(ref['loss'] - plain['loss']) > 1e-3 * abs(plain['loss'])          # the variant is not a no-op
    if kind == 'weighted':
        assert np.array_equal(model._inplane.cpu().numpy(), whw)
    loss = float(model.train_step_device(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                         apply_optimizer=False).item())
    tol = 1e-5 if precision == 'fp32' else 1e-2
    assert abs(loss - ref['loss']) <= tol * abs(ref['loss']), (loss, ref['loss'])
    g = model.grads.cpu().numpy()
    checked = 0
    for (name, is_state, off, shape), rg in zip(model.tensors, ref['grads']):
        if is_state or np.linalg.norm(rg) < 1e-12:
            continue
        last = name.startswith(('head/', 'dec1.conv_b/'))
        if precision == 'bf16' and (not last or name == 'dec1.conv_b/bias'):
            continue          # deep bf16 gradients: conditioning is the subject of test_train_step_matches_oracle
        mine = g[off:off + int(np.prod(shape))].reshape(shape).astype(np.float64)
        rl2 = np.linalg.norm(mine - rg) / np.linalg.norm(rg)
        lim = (3e-3 if not name.endswith('/bias') else 1e-2) if precision == 'fp32' else 7e-2
        assert rl2 <= lim, (name, rl2)
        checked += 1
    assert checked >= (20 if precision == 'fp32' else 4), checked


@pytest.mark.parametrize('kind', ['mse', 'masked', 'weighted', 'bce_dice'])
def test_evaluate_and_dice_metrics_match_oracle(kind):
    """model.evaluate (the validation_data leg of fit): inference-mode loss of every loss kind and the dice_coef*
    metrics from rvip_heat_stats, against the oracle's heat maps reduced in numpy float64."""
    from cmr_landmark_detection_b200.models import Loss_and_metrics as metr
    from oracle import unet_ref as R
    model, cfg, ws, x, y = _setup('fp32', 32, 2, 6, seed=5, randomize_bn=True)
    loss = {'mse': metr.mse, 'masked': metr.loss_with_zero_mask(metr.mse, 0.05),
            'weighted': metr.loss_with_zero_mask(metr.mse, 0.05, weight_inplane=True, xy_shape=32),
            'bce_dice': metr.BceDiceLoss(w_bce=0.7, w_dice=1.3)}[kind]
    mets = [metr.dice_coef, metr.dice_coef_labels, metr.dice_coef_lower, metr.dice_coef_upper, metr.dice_coef_background]
    model.compile(loss=loss, metrics=mets)
    with pytest.raises(ValueError):
        model.compile(metrics=[metr.dice_coef_rv])            # channel -3 of a 2-channel heat map
    model.compile(loss=loss, metrics=mets)
    res = model.evaluate(x, y, batch_size=4, return_dict=True)      # batches of 4 and 2
    ref = R.predict(cfg, ws, x).astype(np.float64)
    t = y.astype(np.float64)
    want = {k: 0.0 for k in res}
    for lo, hi in ((0, 4), (4, 6)):
        p, tt = ref[lo:hi], t[lo:hi]
        hp = torch.from_numpy(p).permute(0, 3, 1, 2)
        ht = torch.from_numpy(tt).permute(0, 3, 1, 2)
        args = dict(mask_smaller_than=0.05) if kind in ('masked', 'weighted') else {}
        if kind == 'bce_dice':
            args = dict(w_bce=0.7, w_dice=1.3)
        lv = float(R.loss_torch(hp, ht, kind, weights_hw=R.inplane_weights(32, 32) if kind == 'weighted' else None, **args))

        def dice(sel):
            return (2 * (tt[..., sel] * p[..., sel]).sum() + 1) / (tt[..., sel].sum() + p[..., sel].sum() + 1)
        vals = {'loss': lv, 'dice_coef': dice([0, 1]), 'dice_coef_labels': dice([0, 1]), 'dice_coef_lower': dice([0]),
                'dice_coef_upper': dice([1]), 'dice_coef_background': dice([0])}
        for k in want:
            want[k] += vals[k] * (hi - lo) / 6.0
    for k in want:
        assert abs(res[k] - want[k]) <= 2e-4 * abs(want[k]) + 1e-6, (k, res[k], want[k])
    assert abs(model.evaluate(x, y, batch_size=4) - want['loss']) <= 2e-4 * abs(want['loss']) + 1e-6


def test_fit_logs_metrics_under_keras_names(tmp_path):
    """train_model.py:54-59, 83, 105-112: metrics passed to create_unet appear in the epoch logs as <name> / val_<name>,
    and MONITOR_FUNCTION / SAVE_MODEL_FUNCTION can select them (get_callbacks)."""
    import os
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models import Loss_and_metrics as metr
    from cmr_landmark_detection_b200.models.Unets import create_unet
    from cmr_landmark_detection_b200.utils.KerasCallbacks import get_callbacks
    config = dict(BASE, DIM=[32, 32], PRECISION='bf16', LEARNING_RATE=2e-3, LOSS_FUNCTION='BcdDiceLoss',
                  MODEL_PATH=str(tmp_path / 'model'), TENSORBOARD_PATH=str(tmp_path / 'tb'),
                  MONITOR_FUNCTION='val_dice_coef_labels', MONITOR_MODE='max', SAVE_MODEL_FUNCTION='val_dice_coef_labels',
                  SAVE_MODEL_MODE='max')
    model = create_unet(config, metrics=[metr.dice_coef_labels, metr.dice_coef_lower, metr.dice_coef_upper])
    x, y = synth.make_batch(10, 32, 32, seed=12)
    h = model.fit(x, y, batch_size=4, epochs=3, callbacks=get_callbacks(config), validation_data=(x[:4], y[:4]),
                  verbose=0, shuffle=False)
    for k in ('loss', 'dice_coef_labels', 'dice_coef_lower', 'dice_coef_upper', 'val_loss', 'val_dice_coef_labels', 'lr'):
        assert k in h.history and len(h.history[k]) == 3 and np.isfinite(h.history[k]).all(), k
    assert all(0.0 < v < 1.0 for v in h.history['dice_coef_labels'])
    # 10 samples in batches of 4 -> ceil = 3 steps per epoch (Keras), the ragged last batch included
    assert model.optimizer.iterations == 9
    assert os.path.exists(os.path.join(config['MODEL_PATH'], 'model.h5')) or \
        os.path.exists(os.path.join(config['MODEL_PATH'], 'model.h5.npz'))


def test_binding_cache_is_bounded_and_targets_are_validated():
    from cmr_landmark_detection_b200.models.Unets import create_unet
    model = create_unet(dict(BASE, DIM=[32, 32], PRECISION='bf16', MAX_BINDINGS=2))
    for b in (1, 2, 3, 4, 2):
        out = model.predict(np.zeros((b, 32, 32, 1), np.float32), batch_size=b)
        assert out.shape == (b, 32, 32, 2)
        assert len(model._bindings) <= 2
    with pytest.raises(ValueError):
        model.train_on_batch(np.zeros((2, 32, 32, 1), np.float32), np.zeros((2, 32, 32, 3), np.float32))
    with pytest.raises(ValueError):
        model.train_step_device(torch.zeros((2, 32, 32, 1), device='cuda'),
                                torch.zeros((2, 32, 32, 2), device='cuda', dtype=torch.float64))
