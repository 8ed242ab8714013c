"""get_optimizer with the reference's signature (src/models/ModelUtils.py:75-118)."""
import logging

from ..runtime.model import SGD, Adam


def get_optimizer(config, name_suff=''):
    """Returns the optimizer create_unet compiles with. Adam (every shipped config,
    exp/template_cfgs/example_config.json:44) and SGD (ModelUtils.py:109-111: lr, nesterov=True, momentum 0) run on the
    device path; the reference reads EPSILON and DECAY but never passes them on (ModelUtils.py:86-87 vs :107), so they
    are ignored here as well."""
    opt = config.get('OPTIMIZER', 'Adam').lower()
    lr = config.get('LEARNING_RATE', 0.001)
    if opt == 'adam':
        optimizer = Adam(lr=lr, name=opt + name_suff)
    elif opt == 'sgd':
        optimizer = SGD(lr=lr, nesterov=True, name=opt + name_suff)
    elif opt in ('adagrad', 'rmsprop', 'adadelta', 'radam', 'nadam'):
        raise NotImplementedError("OPTIMIZER=%r is not implemented on the B200 path (Adam and SGD are)" % opt)
    else:
        optimizer = Adam()        # ModelUtils.py:113-115: unknown name -> Adam with standard parameters
    logging.debug('Optimizer: {}'.format(opt))
    return optimizer
