"""Callbacks for RvipUNet.fit with the semantics of the tf.keras callbacks the reference hands to model.fit
(src/utils/KerasCallbacks.py:20-114 get_callbacks): best-only weight checkpoints, ReduceLROnPlateau, EarlyStopping,
learning-rate logging and the optional polynomial decay.  Pure host logic -- the reference's classes need TensorFlow;
these only touch the model attributes RvipUNet provides (optimizer.lr, stop_training, save_weights).  The image writers
(ImageSaver / CustomImageWritertf2) belong to plotting and are not provided."""
from __future__ import annotations

import logging
import os

import numpy as np


class Callback:
    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass


def _monitor_op(monitor, mode):
    """Keras: mode 'auto' maximises metrics whose name contains 'acc' (or starts with 'fmeasure'), else minimises."""
    if mode not in ('auto', 'min', 'max'):
        mode = 'auto'
    if mode == 'max' or (mode == 'auto' and ('acc' in monitor or monitor.startswith('fmeasure'))):
        return np.greater, -np.inf
    return np.less, np.inf


class ModelCheckpoint(Callback):
    """tf.keras.callbacks.ModelCheckpoint(save_best_only=True, save_weights_only=True, save_freq='epoch') as configured at
    KerasCallbacks.py:54-61."""

    def __init__(self, filepath, monitor='val_loss', verbose=0, save_best_only=False, save_weights_only=True,
                 mode='auto', save_freq='epoch'):
        if not save_weights_only:
            raise NotImplementedError('only save_weights_only=True (what the reference uses) is implemented')
        self.filepath, self.monitor, self.verbose, self.save_best_only = filepath, monitor, verbose, save_best_only
        self.monitor_op, self.best = _monitor_op(monitor, mode)

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        path = self.filepath.format(epoch=epoch + 1, **logs)
        if not self.save_best_only:
            self.model.save_weights(path)
            return
        current = logs.get(self.monitor)
        if current is None:
            logging.warning('Can save best model only with %s available, skipping.', self.monitor)
            return
        if self.monitor_op(current, self.best):
            if self.verbose:
                print('\nEpoch %05d: %s improved from %0.5f to %0.5f, saving model to %s' %
                      (epoch + 1, self.monitor, self.best, current, path))
            self.best = current
            self.model.save_weights(path)
        elif self.verbose:
            print('\nEpoch %05d: %s did not improve from %0.5f' % (epoch + 1, self.monitor, self.best))


class ReduceLROnPlateau(Callback):
    """tf.keras.callbacks.ReduceLROnPlateau (TF 2.3 semantics) as configured at KerasCallbacks.py:63-70."""

    def __init__(self, monitor='val_loss', factor=0.1, patience=10, verbose=0, mode='auto', min_delta=1e-4, cooldown=0,
                 min_lr=0):
        if factor >= 1.0:
            raise ValueError('ReduceLROnPlateau does not support a factor >= 1.0.')
        self.monitor, self.factor, self.patience, self.verbose = monitor, factor, patience, verbose
        self.min_delta, self.cooldown, self.min_lr, self.mode = min_delta, cooldown, min_lr, mode
        self._reset()

    def _reset(self):
        op, self.best = _monitor_op(self.monitor, self.mode)
        if op is np.less:
            self.monitor_op = lambda a, b: np.less(a, b - self.min_delta)
        else:
            self.monitor_op = lambda a, b: np.greater(a, b + self.min_delta)
        self.cooldown_counter = 0
        self.wait = 0

    def on_train_begin(self, logs=None):
        self._reset()

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        logs['lr'] = float(self.model.optimizer.lr)
        current = logs.get(self.monitor)
        if current is None:
            logging.warning('Reduce LR on plateau conditioned on metric `%s` which is not available.', self.monitor)
            return
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        if self.monitor_op(current, self.best):
            self.best = current
            self.wait = 0
        elif not self.cooldown_counter > 0:
            self.wait += 1
            if self.wait >= self.patience:
                old_lr = float(self.model.optimizer.lr)
                if old_lr > self.min_lr:
                    new_lr = max(old_lr * self.factor, self.min_lr)
                    self.model.optimizer.lr = new_lr
                    if self.verbose:
                        print('\nEpoch %05d: ReduceLROnPlateau reducing learning rate to %s.' % (epoch + 1, new_lr))
                    self.cooldown_counter = self.cooldown
                    self.wait = 0


class EarlyStopping(Callback):
    """tf.keras.callbacks.EarlyStopping (TF 2.3 semantics) as configured at KerasCallbacks.py:107-112."""

    def __init__(self, monitor='val_loss', min_delta=0, patience=0, verbose=0, mode='auto', baseline=None,
                 restore_best_weights=False):
        self.monitor, self.patience, self.verbose, self.baseline = monitor, patience, verbose, baseline
        self.restore_best_weights = restore_best_weights
        self.monitor_op, _ = _monitor_op(monitor, mode)
        self.min_delta = abs(min_delta) * (1 if self.monitor_op is np.greater else -1)
        self.stopped_epoch = 0

    def on_train_begin(self, logs=None):
        self.wait = 0
        self.stopped_epoch = 0
        self.best_weights = None
        self.best = self.baseline if self.baseline is not None else (np.inf if self.monitor_op is np.less else -np.inf)

    def on_epoch_end(self, epoch, logs=None):
        current = (logs or {}).get(self.monitor)
        if current is None:
            logging.warning('Early stopping conditioned on metric `%s` which is not available.', self.monitor)
            return
        if self.monitor_op(current - self.min_delta, self.best):
            self.best = current
            self.wait = 0
            if self.restore_best_weights:
                self.best_weights = self.model.get_weights()
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.stopped_epoch = epoch
                self.model.stop_training = True
                if self.restore_best_weights and self.best_weights is not None:
                    self.model.set_weights(self.best_weights)

    def on_train_end(self, logs=None):
        if self.stopped_epoch > 0 and self.verbose:
            print('Epoch %05d: early stopping' % (self.stopped_epoch + 1))


class OptimizerChanger(EarlyStopping):
    """Switches the optimizer instead of ending the run (KerasCallbacks.py:245-278): an EarlyStopping that, when training
    ends, calls `on_train_end(config, train_generator, val_generator, model, metrics, last_epoch)`."""

    def __init__(self, on_train_end, train_generator, val_generator, config, metrics, **kwargs):
        self.do_on_train_end = on_train_end
        self.train_generator, self.val_generator = train_generator, val_generator
        self.config, self.metrics = config, metrics
        self.current_epoch = 0
        super().__init__(**kwargs)

    def on_epoch_end(self, epoch, logs=None):
        super().on_epoch_end(epoch, logs)
        self.current_epoch = epoch

    def on_train_end(self, logs=None):
        super().on_train_end(logs)
        self.do_on_train_end(self.config, self.train_generator, self.val_generator, self.model, self.metrics,
                             self.current_epoch)


def finetune_with_SGD(config, train_g, val_g, model, metrics, epoch_init):
    """Fine-tunes a converged model with plain SGD (KerasCallbacks.py:280-306): recompiles with
    tf.keras.optimizers.SGD(name='SGD') (learning rate 0.01, no momentum) and fits again from `epoch_init` with the
    callbacks of get_callbacks WITHOUT metrics, i.e. with a plain EarlyStopping (no recursion)."""
    from ..runtime.model import SGD
    model.compile(optimizer=SGD(name='SGD'), loss=config.get('LOSS_FUNCTION', None), metrics=metrics)
    return model.fit(x=train_g, epochs=config['EPOCHS'], callbacks=get_callbacks(config, train_g, val_g),
                     steps_per_epoch=len(train_g), validation_data=val_g, max_queue_size=20, initial_epoch=epoch_init,
                     workers=0, verbose=1)


class LRLogger(Callback):
    """Stand-in for LRTensorBoard (KerasCallbacks.py:166-176): records the learning rate of every epoch in the logs and
    appends 'epoch,lr,<logs...>' lines to <log_dir>/lr_log.csv instead of TensorBoard event files."""

    def __init__(self, log_dir=None, **_ignored):
        self.log_dir = log_dir
        self.history = []

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        logs.update({'lr': float(self.model.optimizer.lr)})
        self.history.append((epoch, logs['lr']))
        if self.log_dir:
            os.makedirs(self.log_dir, exist_ok=True)
            with open(os.path.join(self.log_dir, 'lr_log.csv'), 'a') as f:
                f.write('%d,%s\n' % (epoch, ','.join('%s=%.8g' % kv for kv in sorted(logs.items()))))


class PolynomialDecay:
    """KerasCallbacks.py:123-139: alpha = initAlpha * (1 - epoch / maxEpochs) ** power."""

    def __init__(self, maxEpochs=100, initAlpha=0.01, power=1.0):
        self.maxEpochs, self.initAlpha, self.power = maxEpochs, initAlpha, power

    def __call__(self, epoch):
        decay = (1 - (epoch / float(self.maxEpochs))) ** self.power
        return float(self.initAlpha * decay)


class LearningRateScheduler(Callback):
    """tf.keras.callbacks.LearningRateScheduler: lr = schedule(epoch) at the beginning of every epoch."""

    def __init__(self, schedule, verbose=0):
        self.schedule, self.verbose = schedule, verbose

    def on_epoch_begin(self, epoch, logs=None):
        lr = float(self.schedule(epoch))
        self.model.optimizer.lr = lr
        if self.verbose:
            print('\nEpoch %05d: LearningRateScheduler setting learning rate to %s.' % (epoch + 1, lr))

    def on_epoch_end(self, epoch, logs=None):
        if logs is not None:
            logs['lr'] = float(self.model.optimizer.lr)


def get_callbacks(config=None, batch_generator=None, validation_generator=None, metrics=None):
    """Same list, order and config keys as the reference's get_callbacks (KerasCallbacks.py:20-114) for the callbacks
    that drive training: ModelCheckpoint (best only, weights only), ReduceLROnPlateau (cooldown 2), learning-rate log,
    optional polynomial decay, EarlyStopping."""
    config = config or {}
    os.makedirs(config['MODEL_PATH'], exist_ok=True)
    if config.get('SAVE_LEARNING_PROGRESS_AS_PNG', False) or config.get('SAVE_LEARNING_PROGRESS_AS_TF', False):
        logging.warning('image-writer callbacks (plotting) are outside the hot path and are skipped')
    cbs = [ModelCheckpoint(os.path.join(config['MODEL_PATH'], 'model.h5'), verbose=1, save_best_only=True,
                           save_weights_only=True, monitor=config.get('SAVE_MODEL_FUNCTION', 'loss'),
                           mode=config.get('SAVE_MODEL_MODE', 'min'), save_freq='epoch'),
           ReduceLROnPlateau(monitor=config.get('MONITOR_FUNCTION', 'loss'), factor=config.get('DECAY_FACTOR', 0.5),
                             patience=config.get('REDUCE_LR_ON_PLAEAU_PATIENCE', 5), verbose=1, cooldown=2,
                             mode=config.get('MONITOR_MODE', 'auto'), min_lr=config.get('MIN_LR', 1e-12)),
           LRLogger(log_dir=config.get('TENSORBOARD_PATH', 'temp/tf_log'))]
    if config.get('POLY_LR_DECAY', False):
        cbs.append(LearningRateScheduler(PolynomialDecay(maxEpochs=config.get('EPOCHS', 100),
                                                         initAlpha=config.get('LEARNING_RATE', 1e-4), power=2), verbose=1))
    if metrics:   # optimizer is changed to SGD once Adam stops improving (KerasCallbacks.py:89-105)
        logging.info('optimizer will be changed to SGD after adam does not improve any more')
        cbs.append(OptimizerChanger(on_train_end=finetune_with_SGD, train_generator=batch_generator,
                                    val_generator=validation_generator, config=config, metrics=metrics, patience=15,
                                    verbose=1, monitor=config.get('MONITOR_FUNCTION', 'loss'),
                                    mode=config.get('MONITOR_MODE', 'min')))
        return cbs
    cbs.append(EarlyStopping(patience=config.get('EARLY_STOPPING_PATIENCE', 25), verbose=1,
                             monitor=config.get('MONITOR_FUNCTION', 'loss'), mode=config.get('MONITOR_MODE', 'min')))
    return cbs
