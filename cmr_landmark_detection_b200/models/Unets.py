"""create_unet / get_model with the reference's signatures (src/models/Unets.py:61-133, :984-998),
returning the B200-native model object instead of a compiled tf.keras.Model."""
import logging

from ..runtime.model import RvipUNet
from . import ModelUtils as mutils
from .Loss_and_metrics import mse, resolve_loss


def create_unet(config, metrics=None, networkname='unet', single_model=True, supervision=False):
    """Factory for the 2D RVIP heat-map U-Net.
    :param config: key/value pairs, UPPERCASE keys (exp/template_cfgs/example_config.json). Extra key
                   PRECISION ('bf16' default -> tcgen05 tensor cores, 'fp32' -> CUDA-core parity mode).
    :param metrics: dice_coef* descriptors of models/Loss_and_metrics.py (train_model.py:54-59); evaluated on the device
                    per batch and logged under the Keras names (dice_coef_labels, val_dice_coef_labels, ...)
    :param networkname: model name (the head layer is called 'unet' like the reference's, Unets.py:128)
    :param single_model: True -> sigmoid head + compile, as every caller uses it (train_model.py:83)
    :param supervision: deep-supervision branch (Unets.py:840-863) -- off in all callers, not implemented
    Data parallelism: the reference opens tf.distribute.MirroredStrategy here (Unets.py:70-75); the
    equivalent is one process per GPU with torch.distributed initialised before this call (runtime/dist.py).
    """
    if supervision:
        raise NotImplementedError('supervision=True is not used by any caller of the reference and is not implemented')
    if not single_model:
        raise NotImplementedError('single_model=False (stacked models) belongs to the 3D wrappers (out of scope)')
    if len(config.get('DIM', [224, 224])) != 2:
        raise NotImplementedError('only the 2D U-Net of the RVIP path is implemented')
    model = RvipUNet(config, name=networkname)
    loss_f = resolve_loss(config.get('LOSS_FUNCTION', mse))          # config strings: train_model.py:178-184
    model.compile(optimizer=mutils.get_optimizer(config, networkname), loss={'unet': loss_f}, metrics=metrics)
    logging.info('created %s: %d parameters', networkname, model.count_params())
    return model


def get_model(config=dict(), metrics=None):
    """Unets.py:984-998: build by ARCHITECTURE name; only 'unet' exists on this path."""
    arch = config.get('ARCHITECTURE', 'unet').lower()
    if arch != 'unet':
        raise NotImplementedError('ARCHITECTURE=%r' % arch)
    return create_unet(config, metrics)
