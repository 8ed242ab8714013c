"""Two-rank probe (torchrun): do the replicas stay bit-identical through fit() and its ingredients?"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cmr_landmark_detection_b200 import synth
from cmr_landmark_detection_b200.models.Unets import create_unet
from cmr_landmark_detection_b200.runtime import dist as rdist
from cmr_landmark_detection_b200.utils.KerasCallbacks import get_callbacks
rank, local, world = rdist.init_from_env()
torch.cuda.set_device(local)
prec = sys.argv[1]
config = {'DIM': [64, 64], 'DEPTH': 2, 'FILTERS': 32, 'IMG_CHANNELS': 1, 'MASK_CLASSES': 2, 'BATCH_NORMALISATION': True,
          'BN_FIRST': False, 'ACTIVATION': 'relu', 'PAD': 'same', 'DROPOUT_MIN': 0.0, 'DROPOUT_MAX': 0.0,
          'LEARNING_RATE': 5e-3, 'M_POOL': [2, 2], 'F_SIZE': [3, 3], 'SEED': 7, 'PRECISION': prec}


def check(m, what):
    torch.cuda.synchronize()
    full = [torch.zeros_like(m.params) for _ in range(world)]
    torch.distributed.all_gather(full, m.params)
    diff = []
    for name, st, off, shape in m.tensors:
        if st:
            continue
        n = int(np.prod(shape))
        if not torch.equal(full[0][off:off + n], full[1][off:off + n]):
            diff.append(name)
    if rank == 0:
        print(what, 'same' if not diff else 'DIVERGED %d tensors: %s' % (len(diff), diff[:8]), flush=True)


xa, ya = synth.make_batch(24, 64, 64, seed=9)
xv, yv = synth.make_batch(12, 64, 64, seed=10)
m = create_unet(config)
m.fit(xa, ya, batch_size=6, epochs=2, verbose=0)
check(m, 'fit plain')
m = create_unet(config)
m.fit(xa, ya, batch_size=6, epochs=2, verbose=0, validation_data=(xv, yv))
check(m, 'fit + validation')
box = [tempfile.mkdtemp() if rank == 0 else None]
torch.distributed.broadcast_object_list(box, src=0)
fcfg = dict(config, MODEL_PATH=os.path.join(box[0], 'model'), TENSORBOARD_PATH=os.path.join(box[0], 'tb'),
            MONITOR_FUNCTION='val_loss', SAVE_MODEL_FUNCTION='val_loss', REDUCE_LR_ON_PLAEAU_PATIENCE=1, EARLY_STOPPING_PATIENCE=3)
m = create_unet(fcfg)
m.fit(xa, ya, batch_size=6, epochs=6, verbose=0, validation_data=(xv, yv), callbacks=get_callbacks(fcfg))
check(m, 'fit + validation + callbacks')
torch.distributed.barrier(); torch.distributed.destroy_process_group()
