"""pred_fold with the reference's signature (src/models/predict_model.py:7-201): rebuild the model from the config
(:75), load MODEL_PATH/model.h5 (:76), and per (patient, phase) generator: model.predict at batch 1 (:89, :143),
threshold -> label volume (:149-156), optional largest-connected-component filter (:159-161), write gt / pred / cmr
volumes (:174-186).  Heat maps, label maps and the CC filter run on the device; the volumes are written as NRRD by a
small writer (SimpleITK is not installed).  What cannot run here is the reference's file-based generator construction
(get_trainings_files / DataGenerator, :56-58, :133) and undo_generator_steps (:167-172, SimpleITK resampling): the
generators come from the config (PRED_GENERATORS), and DATA_PATH_ORIG-based un-resampling raises a clear error."""
from __future__ import annotations

import logging
import os
from time import time

import numpy as np
import torch

from ..extract import cc_filter_device, label_map_device


def predict_label_volume(model, x: np.ndarray, thr: float = 0.5, batch_size: int = 1) -> np.ndarray:
    """preds = model.predict(x); preds_flat[preds[...,0]>thr]=1; preds_flat[preds[...,1]>thr]=2 -> uint8 [N,H,W]
    (predict_model.py:143-156, :170-171). pred_fold predicts with BATCHSIZE=1 (:89)."""
    preds = model.predict(np.asarray(x, np.float32), batch_size=batch_size)
    heat = torch.from_numpy(preds).to(model.device)
    return label_map_device(heat, thr).cpu().numpy()


def _pred_generators(config: dict):
    """[(patient, phase, Sequence)] -- predict_model.py:101-133 builds one batch-1 DataGenerator per patient and phase
    (ED / ES) from the fold's .nrrd files."""
    gens = config.get('PRED_GENERATORS')
    if gens is None:
        raise RuntimeError(
            'pred_fold: building the per-patient DataGenerators from files (predict_model.py:56-58, 101-133) needs the '
            "reference's SimpleITK data plane, which is outside the B200 hot path. Pass config['PRED_GENERATORS'] = "
            '[(patient, phase, sequence), ...] with batch-1 sequences whose items are (x [1,H,W,1], y [1,H,W,C]).')
    return list(gens)


def pred_fold(config, debug=True):
    from ..utils.nrrd_io import write_nrrd
    from .Unets import create_unet
    t0 = time()
    try:
        model = create_unet(config)
        model.load_weights(os.path.join(config['MODEL_PATH'], 'model.h5'))
        logging.info('loaded model weights as h5 file')
        pred_path = os.path.join(config.get('EXP_PATH'), 'pred')
        gt_path = os.path.join(config.get('EXP_PATH'), 'gt')
        os.makedirs(pred_path, exist_ok=True)
        os.makedirs(gt_path, exist_ok=True)
        if config.get('DATA_PATH_ORIG') and config.get('UNDO_GENERATOR_STEPS', False):
            raise RuntimeError('pred_fold: undo_generator_steps (predict_model.py:164-172: resampling back onto the original '
                               'CMR grid) needs SimpleITK; volumes are written with the config SPACING instead')
        for p, current_phase, validation_generator in _pred_generators(config):
            logging.info('patient: {}, phase: {}, files: {}'.format(p, current_phase, len(validation_generator)))
            items = [validation_generator[i] for i in range(len(validation_generator))]
            gts = np.stack([np.squeeze(y) for x, y in items])                 # :136
            gts_cmr = np.stack([np.squeeze(x) for x, y in items])             # :139
            preds = model.predict(validation_generator)                       # :143, batch 1 per slice
            dev = model.device
            gts_flat = label_map_device(torch.from_numpy(np.ascontiguousarray(gts, dtype=np.float32)).to(dev), 0.5)
            preds_flat = label_map_device(torch.from_numpy(preds).to(dev), 0.5)         # :149-156
            if config.get('CC_FILTER', False):                                # :159-161
                preds_flat = cc_filter_device(preds_flat, 8)
            exp_spacing = tuple(reversed(config.get('SPACING', (1.0, 1.0)))) + (10,)    # :179-180
            write_nrrd(os.path.join(gt_path, '{}_{}_msk.nrrd'.format(p, current_phase)), gts_flat.cpu().numpy(), exp_spacing)
            write_nrrd(os.path.join(pred_path, '{}_{}_msk.nrrd'.format(p, current_phase)), preds_flat.cpu().numpy(),
                       exp_spacing)
            write_nrrd(os.path.join(pred_path, '{}_{}_cmr.nrrd'.format(p, current_phase)),
                       np.ascontiguousarray(gts_cmr, dtype=np.float32), exp_spacing)
        logging.info('done! Check the folders \n{} and \n{} for files'.format(gt_path, pred_path))
    except Exception as e:
        if debug:
            raise           # the reference logs and swallows (:192-193); debug=True (its default) surfaces the error here
        logging.error(e)
    logging.info('pred on fold {} finished after {:0.3f} sec'.format(config.get('FOLD'), time() - t0))
    return True
