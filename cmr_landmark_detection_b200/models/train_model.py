"""train_fold with the reference's signature (src/models/train_model.py:1-132): the compute steps -- create_unet
(:83) -> model.summary (:84-89) -> get_callbacks (:102) -> model.fit (:105-112) -> pred_fold (:123-124) -- run on the
B200 path over any keras.utils.Sequence-like generator.  The one step that cannot run here is the reference's own data
plane (get_trainings_files + DataGenerator, :68-78: SimpleITK / albumentations file I/O, SURVEY section 2 rows 9-12, out
of scope): a caller hands the generators over in the config instead (TRAIN_GENERATOR / VAL_GENERATOR); without them a
clear error names that step."""
from __future__ import annotations

import json
import logging
import os
from time import time


def _fold_paths(config: dict) -> dict:
    """train_model.py:31-48: per-fold sub-folders under EXP_PATH."""
    exp_path = config.get('EXP_PATH')
    if exp_path is None:
        raise KeyError("train_fold: config['EXP_PATH'] is required (train_model.py:40)")
    fold = config.get('FOLD', 0)
    fold_path = os.path.join(exp_path, 'f{}'.format(fold))
    out = {'EXPERIMENT': '{}f{}'.format(config.get('EXPERIMENT'), fold), 'FOLD_PATH': fold_path,
           'MODEL_PATH': os.path.join(fold_path, 'model'), 'TENSORBOARD_PATH': os.path.join(fold_path, 'tensorboard_logs'),
           'CONFIG_PATH': os.path.join(fold_path, 'config')}
    for k in ('MODEL_PATH', 'TENSORBOARD_PATH', 'CONFIG_PATH'):
        os.makedirs(out[k], exist_ok=True)
    return out


def _generators(config: dict, in_memory: bool):
    """train_model.py:68-78.  The reference builds DataGenerator objects from .nrrd files; that data plane needs
    SimpleITK / albumentations and is out of scope, so ready-made Sequence objects are taken from the config."""
    tr, va = config.get('TRAIN_GENERATOR'), config.get('VAL_GENERATOR')
    if tr is not None:
        return tr, va
    try:
        from src.data.Dataset import get_trainings_files        # the reference's own module, if a caller provides it
        from src.data.Generators import DataGenerator
    except Exception as e:
        raise RuntimeError(
            'train_fold: the file-based data plane of the reference (get_trainings_files + DataGenerator, '
            'train_model.py:68-78) needs SimpleITK / albumentations, which are not part of the B200 hot path. Pass '
            "keras.utils.Sequence-like objects as config['TRAIN_GENERATOR'] / config['VAL_GENERATOR'] "
            '(items -> (x float32 [B,H,W,1], y float32 [B,H,W,C])). Import error: %s' % (e,)) from e
    x_train, y_train, x_val, y_val = get_trainings_files(data_path=config.get('DATA_PATH_SAX'),
                                                         path_to_folds_df=config.get('DF_FOLDS'), fold=config.get('FOLD'))
    val_config = dict(config, AUGMENT_GRID=False, AUGMENT=False, HIST_MATCHING=False)
    return (DataGenerator(x_train, y_train, config=config, in_memory=in_memory),
            DataGenerator(x_val, y_val, config=val_config, in_memory=in_memory))


def train_fold(config, in_memory=True):
    """Trains one cross-validation fold and predicts its hold-out set; returns True like the reference."""
    from ..utils.KerasCallbacks import get_callbacks
    from . import Loss_and_metrics as metr
    from . import Unets as modelmanager
    t0 = time()
    config = dict(config)
    config.update(_fold_paths(config))
    epochs = config.get('EPOCHS', 100)
    # train_model.py:54-59; a metric whose channel the heat map does not have (dice_coef_rv = channel -3 of the two RVIP
    # channels) fails inside TensorFlow too -- the authors' own runs log labels / lower / upper (Train_tests.ipynb cell 11)
    metrics = []
    for m in (metr.dice_coef_labels, metr.dice_coef_myo, metr.dice_coef_lv, metr.dice_coef_rv):
        try:
            m.rvip_channels(int(config.get('MASK_CLASSES', 3)))
            metrics.append(m)
        except ValueError as e:
            logging.warning('metric skipped: %s', e)
    # train_model.py:62 init_config: persist the UPPERCASE, serialisable keys
    def _plain(v):
        try:
            json.dumps(v)
            return True
        except TypeError:
            return False
    with open(os.path.join(config['CONFIG_PATH'], 'config.json'), 'w') as fh:
        json.dump({k: v for k, v in config.items() if k.isupper() and _plain(v)}, fh, indent=1)
    batch_generator, validation_generator = _generators(config, in_memory)
    logging.info('Create model')
    model = modelmanager.create_unet(config, metrics, supervision=False)
    model.summary()
    with open(os.path.join(config['EXP_PATH'], 'model_summary.txt'), 'w') as fh:
        model.summary(print_fn=lambda x: fh.write(x + '\n'))
    initial_epoch = 0
    cb = get_callbacks(config, batch_generator, validation_generator)
    model.fit(x=batch_generator, validation_data=validation_generator, epochs=epochs, callbacks=cb,
              initial_epoch=initial_epoch, max_queue_size=config.get('QUEUE_SIZE', 12), verbose=config.get('VERBOSE', 1))
    try:
        del model, cb
        from .predict_model import pred_fold
        pred_fold(config)
    except Exception as e:          # train_model.py:128-129 logs and carries on
        logging.error(e)
    logging.info('Fold {} finished after {:0.3f} sec'.format(config.get('FOLD'), time() - t0))
    return True
