"""CPU: the fit() callbacks against hand-traced tf.keras (TF 2.3) semantics on synthetic loss curves."""
import os

import numpy as np

from cmr_landmark_detection_b200.utils import KerasCallbacks as K


class _Opt:
    lr = 1e-3


class _Model:
    def __init__(self):
        self.optimizer = _Opt()
        self.stop_training = False
        self.saved = []

    def save_weights(self, path):
        self.saved.append(path)


def _run(cbs, losses):
    m = _Model()
    for cb in cbs:
        cb.set_model(m)
        cb.on_train_begin({})
    lrs = []
    for e, l in enumerate(losses):
        for cb in cbs:
            cb.on_epoch_begin(e, {})
        logs = {'loss': l}
        for cb in cbs:
            cb.on_epoch_end(e, logs)
        lrs.append(m.optimizer.lr)
        if m.stop_training:
            break
    return m, lrs, e


def test_reduce_lr_on_plateau_patience_cooldown_min_lr():
    # improvement needs > min_delta (1e-4); patience 2, cooldown 2, factor 0.5
    losses = [1.0, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9]
    cb = K.ReduceLROnPlateau(monitor='loss', factor=0.5, patience=2, cooldown=2, min_lr=3e-4)
    _, lrs, _ = _run([cb], losses)
    # epochs 2,3 are the two non-improving waits -> reduce after epoch 3 and start a 2-epoch cooldown; Keras decrements
    # the counter BEFORE testing it, so epoch 4 is swallowed, epoch 5 already counts (wait 1), epoch 6 is wait 2 ->
    # second reduction, clipped at min_lr
    assert np.allclose(lrs, [1e-3, 1e-3, 1e-3, 5e-4, 5e-4, 5e-4, 3e-4, 3e-4, 3e-4])


def test_early_stopping_and_checkpoint_best_only(tmp_path):
    losses = [1.0, 0.8, 0.85, 0.79, 0.9, 0.9, 0.9, 0.9]
    es = K.EarlyStopping(monitor='loss', patience=3, mode='min')
    ck = K.ModelCheckpoint(str(tmp_path / 'model.h5'), monitor='loss', save_best_only=True, save_weights_only=True, mode='min')
    m, _, last = _run([ck, es], losses)
    assert last == 6 and m.stop_training and es.stopped_epoch == 6      # best at epoch 3, three waits: 4, 5, 6
    assert len(m.saved) == 3                                            # epochs 0, 1, 3 improved


def test_get_callbacks_matches_reference_configuration(tmp_path):
    cfg = {'MODEL_PATH': str(tmp_path / 'model'), 'TENSORBOARD_PATH': str(tmp_path / 'tb'), 'MONITOR_FUNCTION': 'loss',
           'POLY_LR_DECAY': True, 'EPOCHS': 10, 'LEARNING_RATE': 1e-3}
    cbs = K.get_callbacks(cfg)
    kinds = [type(c).__name__ for c in cbs]
    assert kinds == ['ModelCheckpoint', 'ReduceLROnPlateau', 'LRLogger', 'LearningRateScheduler', 'EarlyStopping']
    assert cbs[1].cooldown == 2 and cbs[1].factor == 0.5 and cbs[1].patience == 5 and cbs[4].patience == 25
    m, lrs, _ = _run(cbs, [1.0, 0.9, 0.8])
    assert np.allclose(lrs, [1e-3, 1e-3 * 0.81, 1e-3 * 0.64])           # polynomial decay, power 2
    assert os.path.exists(os.path.join(cfg['TENSORBOARD_PATH'], 'lr_log.csv'))


def test_optimizer_changer_switches_to_sgd(tmp_path):
    """get_callbacks(metrics=...) ends with the OptimizerChanger (KerasCallbacks.py:89-105): an EarlyStopping that hands the
    model to finetune_with_SGD when training ends; get_optimizer('sgd') builds SGD(lr, nesterov=True) (ModelUtils.py:109)."""
    from cmr_landmark_detection_b200.models.ModelUtils import get_optimizer
    from cmr_landmark_detection_b200.runtime.model import SGD
    from cmr_landmark_detection_b200.utils import KerasCallbacks as K
    opt = get_optimizer({'OPTIMIZER': 'SGD', 'LEARNING_RATE': 0.02})
    assert isinstance(opt, SGD) and opt.lr == 0.02 and opt.nesterov and opt.momentum == 0.0
    cfg = {'MODEL_PATH': str(tmp_path), 'EPOCHS': 7, 'MONITOR_FUNCTION': 'loss', 'MONITOR_MODE': 'min'}
    cbs = K.get_callbacks(cfg, batch_generator=[1, 2, 3], validation_generator=[4], metrics=['m'])
    ch = cbs[-1]
    assert isinstance(ch, K.OptimizerChanger) and ch.patience == 15 and ch.do_on_train_end is K.finetune_with_SGD
    assert not any(type(c) is K.EarlyStopping for c in cbs)
    assert type(K.get_callbacks(cfg)[-1]) is K.EarlyStopping

    calls = []

    class FakeModel:
        stop_training = False

        def compile(self, optimizer=None, loss=None, metrics=None):
            calls.append(('compile', optimizer, metrics))

        def fit(self, **kw):
            calls.append(('fit', kw))
            return 'history'

    m = FakeModel()
    ch = K.OptimizerChanger(on_train_end=K.finetune_with_SGD, train_generator=[1, 2, 3], val_generator=[4], config=cfg,
                            metrics=['m'], patience=2, monitor='loss', mode='min')
    ch.set_model(m)
    ch.on_train_begin()
    for epoch, loss in enumerate([1.0, 0.9, 0.95, 0.97]):
        ch.on_epoch_end(epoch, {'loss': loss})
    assert m.stop_training and ch.stopped_epoch == 3
    ch.on_train_end()
    (c0, opt, metrics), (c1, kw) = calls
    assert c0 == 'compile' and isinstance(opt, SGD) and opt.lr == 0.01 and opt.momentum == 0.0 and metrics == ['m']
    assert c1 == 'fit' and kw['initial_epoch'] == 3 and kw['epochs'] == 7 and kw['steps_per_epoch'] == 3
    assert type(kw['callbacks'][-1]) is K.EarlyStopping            # no recursion: the second fit stops for good
