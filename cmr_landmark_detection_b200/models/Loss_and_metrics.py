"""Loss descriptors with the names of src/models/Loss_and_metrics.py. The arithmetic runs fused with the
head in csrc/head_loss.cu; these objects only select it (compile(loss=...))."""


class _DeviceLoss:
    def __init__(self, kind, **args):
        self.rvip_kind = kind
        self.rvip_args = args
        self.__name__ = kind

    def __call__(self, y_true, y_pred):
        raise RuntimeError('device losses are evaluated inside rvip_train_step, not called from Python')

    def __repr__(self):
        return '<device loss %s %r>' % (self.rvip_kind, self.rvip_args)


#: tf.keras.losses.mse as imported at Loss_and_metrics.py:6
mse = _DeviceLoss('mse')
MSE = mse


class BceDiceLoss(_DeviceLoss):
    """Loss_and_metrics.py:208-228: w_bce * binary_crossentropy(y_true, y_pred) - w_dice * dice_coef(y_true, y_pred);
    the loss train_model.py:178 selects for LOSS_FUNCTION = 'BcdDiceLoss'."""

    def __init__(self, w_bce=1., w_dice=1., binary=True, name='BcdDiceLoss'):
        if not binary:
            raise NotImplementedError('categorical cross-entropy variant is not implemented')
        super().__init__('bce_dice', w_bce=float(w_bce), w_dice=float(w_dice))
        self.name = '{}_w_{}_{}'.format(name, w_bce, w_dice)


#: Loss_and_metrics.py:231-245 (function form, w_bce = 0.5)
bce_dice_loss = BceDiceLoss(w_bce=0.5, w_dice=1.)


def loss_with_zero_mask(loss=mse, mask_smaller_than=0.01, weight_inplane=False, xy_shape=224):
    """Loss_and_metrics.py:40-89: `loss` only where y_true > mask_smaller_than, optionally times the concentric
    in-plane ramp (+ K.epsilon()). The reference squeezes the mask on axis -1 (needs C == 1); for the two RVIP
    channels the per-pixel mask is any_c(y_true > thr) -- an extension that equals the reference for C == 1."""
    if getattr(loss, 'rvip_kind', None) != 'mse':
        raise NotImplementedError('loss_with_zero_mask is implemented for loss=mse')
    return _DeviceLoss('weighted' if weight_inplane else 'masked', mask_smaller_than=float(mask_smaller_than),
                       xy_shape=int(xy_shape))


def resolve_loss(loss):
    """One place that turns whatever a caller hands over as a loss into a device-loss descriptor: a descriptor itself, a
    Keras-style {'unet': loss} dict (Unets.py:130), or a config string -- train_model.py:178-184 selects BceDiceLoss by
    the substring 'BcdDiceLoss' (its spelling) and falls back to MSE otherwise; 'mse' / 'mean_squared_error' are the
    Keras string forms.  Returns None for None (keep the compiled loss)."""
    if isinstance(loss, dict):
        loss = loss.get('unet', next(iter(loss.values())))
    if loss is None or hasattr(loss, 'rvip_kind'):
        return loss
    if isinstance(loss, str):
        low = loss.lower()
        if 'bcddiceloss' in low or 'bcedice' in low.replace('_', '') or low == 'bce_dice':
            return BceDiceLoss()
        if low in ('mse', 'mean_squared_error') or 'mse' in low:
            return mse
        if low in ('masked', 'weighted'):
            return loss_with_zero_mask(weight_inplane=low == 'weighted')
    raise NotImplementedError('loss %r is not implemented on the device path (MSE / masked / weighted MSE and '
                              'BCE+Dice are)' % (loss,))


class _DiceMetric:
    """dice_coef restricted to a channel selection (Loss_and_metrics.py:124-171): smooth = 1,
    (2 sum(t p) + 1) / (sum t + sum p + 1) over the WHOLE batch of the selected channels.  Evaluated from the three
    per-channel sums rvip_heat_stats reduces on the device; Keras logs the mean of the per-batch values under the
    function's name (and val_<name>), which is what MONITOR_FUNCTION / SAVE_MODEL_FUNCTION select."""

    def __init__(self, name, select):
        self.__name__ = name
        self.name = name
        self._select = select

    def rvip_channels(self, n_classes):
        idx = self._select(n_classes)
        if isinstance(idx, int):
            if not -n_classes <= idx < n_classes:
                # tf: "slice index -3 of dimension 3 out of bounds" when the heat map has fewer channels
                raise ValueError('%s selects channel %d of a %d-channel output' % (self.__name__, idx, n_classes))
            idx = [idx % n_classes]
        return list(idx)

    def __call__(self, y_true, y_pred):
        raise RuntimeError('device metrics are evaluated by rvip_heat_stats, not called from Python')


dice_coef = _DiceMetric('dice_coef', lambda c: range(c))                               # :165-171
dice_coef_labels = _DiceMetric('dice_coef_labels', lambda c: range(max(c - 3, 0), c))    # :154-161 y[..., -3:]
dice_coef_background = _DiceMetric('dice_coef_background', lambda c: 0)                  # :124-127
dice_coef_rv = _DiceMetric('dice_coef_rv', lambda c: -3)                                 # :129-132
dice_coef_lower = _DiceMetric('dice_coef_lower', lambda c: -2)                           # :134-137
dice_coef_upper = _DiceMetric('dice_coef_upper', lambda c: -1)                           # :139-142
dice_coef_myo = _DiceMetric('dice_coef_myo', lambda c: -2)                               # :144-147
dice_coef_lv = _DiceMetric('dice_coef_lv', lambda c: -1)                                 # :149-152
