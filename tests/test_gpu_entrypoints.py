"""train_fold / pred_fold (src/models/train_model.py:1-132, src/models/predict_model.py:7-201) end to end on the device
with a synthetic in-memory generator standing in for the reference's DataGenerator."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _Gen:
    """keras.utils.Sequence protocol of src/data/Generators.py:136-173: len / getitem -> (x, y) / on_epoch_end."""

    def __init__(self, x, y, bs):
        self.x, self.y, self.bs = x, y, bs
        self.epochs_seen = 0

    def __len__(self):
        return len(self.x) // self.bs

    def __getitem__(self, i):
        return self.x[i * self.bs:(i + 1) * self.bs], self.y[i * self.bs:(i + 1) * self.bs]

    def on_epoch_end(self):
        self.epochs_seen += 1


def test_train_fold_then_pred_fold(tmp_path):
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.evaluate_cv import get_ip_from_rvip_file, get_ip_from_rvip_mask_3d
    from cmr_landmark_detection_b200.utils.nrrd_io import read_nrrd
    from src.models.predict_model import pred_fold, predict_label_volume
    from src.models.train_model import train_fold
    from src.models.Unets import create_unet
    x, y = synth.make_batch(16, 64, 64, seed=3)
    xv, yv = synth.make_batch(8, 64, 64, seed=4)
    config = {'EXP_PATH': str(tmp_path), 'EXPERIMENT': 'synthetic', 'FOLD': 0, 'DIM': [64, 64], 'DEPTH': 2, 'FILTERS': 32,
              'IMG_CHANNELS': 1, 'MASK_CLASSES': 2, 'BATCH_NORMALISATION': True, 'BN_FIRST': False, 'ACTIVATION': 'relu',
              'PAD': 'same', 'DROPOUT_MIN': 0.1, 'DROPOUT_MAX': 0.2, 'LEARNING_RATE': 2e-3, 'M_POOL': [2, 2],
              'F_SIZE': [3, 3], 'SEED': 1, 'EPOCHS': 3, 'BATCHSIZE': 4, 'LOSS_FUNCTION': 'BcdDiceLoss', 'SPACING': [1.2, 1.2],
              'MONITOR_FUNCTION': 'loss', 'SAVE_MODEL_FUNCTION': 'loss', 'VERBOSE': 0, 'CC_FILTER': True,
              'TRAIN_GENERATOR': _Gen(x, y, 4), 'VAL_GENERATOR': _Gen(xv, yv, 4),
              'PRED_GENERATORS': [('patient001', 'ED', _Gen(xv[:4], yv[:4], 1)), ('patient001', 'ES', _Gen(xv[4:], yv[4:], 1))]}
    assert train_fold(config) is True
    fold = tmp_path / 'f0'
    assert (fold / 'model' / 'model.h5').exists() or (fold / 'model' / 'model.h5.npz').exists()
    assert 'Total params' in (tmp_path / 'model_summary.txt').read_text()
    saved = json.load(open(fold / 'config' / 'config.json'))
    assert saved['DEPTH'] == 2 and 'TRAIN_GENERATOR' not in saved
    assert config['TRAIN_GENERATOR'].epochs_seen == 3
    # train_fold ran pred_fold on the fold's hold-out generators (train_model.py:123-124)
    for ph in ('ED', 'ES'):
        for f in (tmp_path / 'gt' / ('patient001_%s_msk.nrrd' % ph), tmp_path / 'pred' / ('patient001_%s_msk.nrrd' % ph),
                  tmp_path / 'pred' / ('patient001_%s_cmr.nrrd' % ph)):
            assert f.exists(), f
    # the written prediction == the same steps by hand: rebuild, load the checkpoint, predict at batch 1, threshold, CC
    cfg2 = dict(config, MODEL_PATH=str(fold / 'model'))
    m = create_unet(cfg2)
    m.load_weights(os.path.join(cfg2['MODEL_PATH'], 'model.h5'))
    from cmr_landmark_detection_b200.extract import cc_filter_device
    want = predict_label_volume(m, xv[:4], batch_size=1)
    want = cc_filter_device(torch.from_numpy(want).cuda(), 8).cpu().numpy()
    vol, spacing = read_nrrd(str(tmp_path / 'pred' / 'patient001_ED_msk.nrrd'))
    assert vol.dtype == np.uint8 and vol.shape == (4, 64, 64) and np.array_equal(vol, want)
    assert np.allclose(spacing, (1.2, 1.2, 10))
    gt, _ = read_nrrd(str(tmp_path / 'gt' / 'patient001_ED_msk.nrrd'))
    ref_gt = np.zeros((4, 64, 64), np.uint8)
    ref_gt[yv[:4, ..., 0] > 0.5] = 1
    ref_gt[yv[:4, ..., 1] > 0.5] = 2
    assert np.array_equal(gt, ref_gt)
    cmr, _ = read_nrrd(str(tmp_path / 'pred' / 'patient001_ED_cmr.nrrd'))
    assert np.array_equal(cmr, xv[:4, ..., 0])
    # evaluate_cv.py:385-387 reads those files back
    a1, b1 = get_ip_from_rvip_file(str(tmp_path / 'gt' / 'patient001_ED_msk.nrrd'), keepdim=True)
    a2, b2 = get_ip_from_rvip_mask_3d(ref_gt, keepdim=True)
    assert repr(a1) == repr(a2) and repr(b1) == repr(b2)
    # pred_fold on its own, with the reference's call shape
    assert pred_fold(cfg2, debug=True) is True
    with pytest.raises(RuntimeError, match='PRED_GENERATORS'):
        pred_fold({k: v for k, v in cfg2.items() if k != 'PRED_GENERATORS'})
