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
