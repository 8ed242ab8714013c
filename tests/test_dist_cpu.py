"""Host-side data-parallel logic on CPU (gloo, world_size 2): sharding, bucketed gradient all-reduce and the
equivalence 'N ranks x per-rank batch with per-replica BN == oracle MirroredStrategy semantics'."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cmr_landmark_detection_b200.runtime import dist as rdist
from oracle import unet_ref as R

CFG = {'DIM': [16, 16], 'DEPTH': 2, 'FILTERS': 4, 'IMG_CHANNELS': 1, 'MASK_CLASSES': 2, 'BATCH_NORMALISATION': True,
       'DROPOUT_MIN': 0.0, 'DROPOUT_MAX': 0.0}


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, x, t, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, l, w = rdist.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    cfg = R.cfg_from_config(CFG)
    ws = R.init_weights(cfg, seed=2)
    lo, hi = rdist.shard_range(x.shape[0], rank, world)
    g = R.train_grads(cfg, ws, x[lo:hi], t[lo:hi])
    # flat gradient buffer in parameter order (trainable tensors only), like the device path's `grads`
    tm = R.trainable_mask(cfg)
    flat = torch.from_numpy(np.concatenate([gr.ravel() for gr, m in zip(g['grads'], tm) if m]))
    sizes = [gr.size for gr, m in zip(g['grads'], tm) if m]
    # bucket boundaries at conv-layer starts (kernel tensors), like rvip_abi.cu build_plan
    names = [n for (n, _), m in zip(R.weight_shapes(cfg), tm) if m]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    layer_offsets = [int(offs[i]) for i, n in enumerate(names) if n.endswith('/kernel') and not n.startswith('head')]
    ranges = rdist.bucket_ranges(layer_offsets, int(offs[-1]), n_buckets=4)
    assert sum(c for _, c in ranges) == int(offs[-1]) and ranges[-1][0] == 0
    assert all(ranges[i][0] == ranges[i + 1][0] + ranges[i + 1][1] for i in range(len(ranges) - 1))
    dp = rdist.DataParallel(torch.device('cpu'))
    assert dp.world == world and dp.rank == rank
    dp.allreduce_flat(flat, ranges)
    flat /= world                       # the device path folds 1/world into Adam (grad_scale)
    if rank == 0:
        out['flat'] = flat.numpy().copy()
        out['max'] = dp.max_float(float(rank))
    else:
        dp.max_float(float(rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucketed_allreduce_matches_mirrored_strategy_oracle():
    rng = np.random.default_rng(0)
    x = rng.random((4, 16, 16, 1)).astype(np.float32)
    t = rng.random((4, 16, 16, 2)).astype(np.float32)
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, x, t, out), nprocs=2, join=True)
    cfg = R.cfg_from_config(CFG)
    ws = R.init_weights(cfg, seed=2)
    ref = R.data_parallel_grads(cfg, ws, x, t, world=2)
    tm = R.trainable_mask(cfg)
    want = np.concatenate([g.ravel() for g, m in zip(ref['grads'], tm) if m])
    assert np.allclose(out['flat'], want, rtol=1e-5, atol=1e-8)
    assert out['max'] == 1.0


def test_shard_range_and_bucket_ranges():
    assert rdist.shard_range(64, 1, 2) == (32, 64)
    try:
        rdist.shard_range(10, 0, 4)
        assert False
    except ValueError:
        pass
    r = rdist.bucket_ranges([0, 10, 30, 70], 100, n_buckets=4)
    assert r == [(70, 30), (30, 40), (0, 30)]
