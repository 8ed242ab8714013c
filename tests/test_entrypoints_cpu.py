"""CPU-side checks of the reference-signature entry points and their host helpers (no GPU): loss / metric selection by
the reference's config strings and names, the NRRD container pred_fold writes, and the error a caller gets at the one
step that is out of scope (the reference's SimpleITK data plane)."""
import inspect
import os

import numpy as np
import pytest


def test_signatures_match_reference():
    from src.models.predict_model import pred_fold
    from src.models.train_model import train_fold
    from src.models.Unets import create_unet
    assert str(inspect.signature(train_fold)) == '(config, in_memory=True)'            # train_model.py:1
    assert str(inspect.signature(pred_fold)) == '(config, debug=True)'                 # predict_model.py:7
    assert list(inspect.signature(create_unet).parameters) == ['config', 'metrics', 'networkname', 'single_model',
                                                               'supervision']          # Unets.py:61


def test_resolve_loss_follows_train_model_substring_rule():
    from cmr_landmark_detection_b200.models import Loss_and_metrics as metr
    assert metr.resolve_loss('BcdDiceLoss').rvip_kind == 'bce_dice'                    # train_model.py:178
    assert metr.resolve_loss('BcdDiceLoss_w_1.0_1.0').rvip_kind == 'bce_dice'
    assert metr.resolve_loss('mse') is metr.mse and metr.resolve_loss('mean_squared_error') is metr.mse
    assert metr.resolve_loss({'unet': metr.bce_dice_loss}) is metr.bce_dice_loss       # Unets.py:130
    assert metr.resolve_loss(None) is None
    assert metr.resolve_loss(metr.loss_with_zero_mask(metr.mse, 0.02, True)).rvip_kind == 'weighted'
    with pytest.raises(NotImplementedError):
        metr.resolve_loss('categorical_crossentropy')


def test_dice_metric_channel_selection():
    from cmr_landmark_detection_b200.models import Loss_and_metrics as metr
    assert metr.dice_coef_labels.rvip_channels(2) == [0, 1] and metr.dice_coef_labels.rvip_channels(4) == [1, 2, 3]
    assert metr.dice_coef_lower.rvip_channels(2) == [0] and metr.dice_coef_upper.rvip_channels(2) == [1]
    assert metr.dice_coef_myo.rvip_channels(4) == [2] and metr.dice_coef_lv.rvip_channels(4) == [3]
    assert metr.dice_coef_rv.rvip_channels(3) == [0] and metr.dice_coef_background.rvip_channels(3) == [0]
    with pytest.raises(ValueError):
        metr.dice_coef_rv.rvip_channels(2)
    assert metr.dice_coef_labels.__name__ == 'dice_coef_labels'          # the key Keras logs it under


def test_nrrd_round_trip(tmp_path):
    from cmr_landmark_detection_b200.utils.nrrd_io import read_nrrd, write_nrrd
    rng = np.random.default_rng(0)
    for dt in (np.uint8, np.float32, np.int16):
        v = (rng.random((5, 7, 9)) * 3).astype(dt)
        p = str(tmp_path / ('v_%s.nrrd' % np.dtype(dt).name))
        write_nrrd(p, v, (1.2, 1.5, 10))
        w, sp = read_nrrd(p)
        assert np.array_equal(v, w) and w.dtype == dt and np.allclose(sp, (1.2, 1.5, 10))
    head = open(p, 'rb').read(200).decode('ascii', 'replace')
    assert head.startswith('NRRD0004\n') and 'sizes: 9 7 5' in head and 'encoding: raw' in head
    with pytest.raises(ValueError):
        write_nrrd(p, np.zeros((3, 3)), (1, 1, 1))


def test_train_fold_names_the_out_of_scope_step(tmp_path):
    from src.models.train_model import train_fold
    cfg = {'EXP_PATH': str(tmp_path), 'EXPERIMENT': 'e', 'FOLD': 0, 'MASK_CLASSES': 2}
    with pytest.raises(RuntimeError, match='TRAIN_GENERATOR'):
        train_fold(cfg)
    assert os.path.isdir(tmp_path / 'f0' / 'model') and os.path.exists(tmp_path / 'f0' / 'config' / 'config.json')
    with pytest.raises(KeyError):
        train_fold({})
