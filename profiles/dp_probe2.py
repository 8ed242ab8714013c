import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from cmr_landmark_detection_b200 import synth
from cmr_landmark_detection_b200.models.Unets import create_unet
from cmr_landmark_detection_b200.runtime import dist as rdist
rank, local, world = rdist.init_from_env()
torch.cuda.set_device(local)
config = {'DIM': [64, 64], 'DEPTH': 2, 'FILTERS': 32, 'IMG_CHANNELS': 1, 'MASK_CLASSES': 2, 'BATCH_NORMALISATION': True,
          'BN_FIRST': False, 'ACTIVATION': 'relu', 'PAD': 'same', 'DROPOUT_MIN': 0.0, 'DROPOUT_MAX': 0.0,
          'LEARNING_RATE': 5e-3, 'M_POOL': [2, 2], 'F_SIZE': [3, 3], 'SEED': 7, 'PRECISION': 'fp32'}
def same(m):
    torch.cuda.synchronize()
    full = [torch.zeros_like(m.params) for _ in range(world)]
    torch.distributed.all_gather(full, m.params)
    return bool(torch.equal(full[0], full[1]))
xa, ya = synth.make_batch(24, 64, 64, seed=9)
xv, yv = synth.make_batch(12, 64, 64, seed=10)
m = create_unet(config)
xs = [torch.from_numpy(xa[6*i+3*rank:6*i+3*rank+3]).cuda() for i in range(4)]
ys = [torch.from_numpy(ya[6*i+3*rank:6*i+3*rank+3]).cuda() for i in range(4)]
log = []
for ep in range(3):
    for i in range(4):
        m.train_step_device(xs[i], ys[i])
        log.append(('e%d s%d' % (ep, i), same(m)))
    v = m.evaluate(xv, yv, batch_size=6)
    log.append(('e%d eval' % ep, same(m)))
if rank == 0:
    print(os.environ.get('RVIP_NO_INLINE_ADAM'), log, flush=True)
torch.distributed.barrier(); torch.distributed.destroy_process_group()
