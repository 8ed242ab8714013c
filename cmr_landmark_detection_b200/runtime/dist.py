"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch on the
GPU box, gloo in CPU tests).  Mirrors tf.distribute.MirroredStrategy as create_unet uses it
(src/models/Unets.py:70-75): replicas hold identical weights, BatchNorm statistics stay per
replica, gradients are all-reduced every step.  The data path has exactly one exchange step, the
gradient all-reduce; it is issued per bucket on a side stream as soon as backward has produced the
bucket (CUDA events recorded inside rvip_train_step), so it overlaps the remaining dgrad / wgrad."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """Initialises torch.distributed from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, local_rank, world). No-op for single-process runs."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank r owns samples [r*per, (r+1)*per) of the global batch (BATCHSIZE is global, Unets.py:70-75)."""
    if n_items % world != 0:
        raise ValueError('global batch %d is not divisible by %d replicas' % (n_items, world))
    per = n_items // world
    return rank * per, (rank + 1) * per


def bucket_ranges(layer_offsets: Sequence[int], total: int, n_buckets: int = 4) -> List[Tuple[int, int]]:
    """Host restatement of the C plan's bucketing (rvip_abi.cu build_plan): contiguous suffixes of the flat
    gradient buffer in backward-completion order, closed whenever they reach total / n_buckets."""
    target = max(1, total // n_buckets)
    out, end = [], total
    for i in range(len(layer_offsets) - 1, -1, -1):
        begin = layer_offsets[i]
        if end - begin >= target or i == 0:
            out.append((begin, end - begin))
            end = begin
    return out


class DataParallel:
    def __init__(self, device, enabled: bool = True):
        self.device = device
        self.enabled = bool(enabled) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.enabled else 1
        self.rank = dist.get_rank() if self.enabled else 0
        self._comm_stream = None

    def comm_stream(self):
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        return self._comm_stream

    def allreduce_buckets(self, grads: torch.Tensor, ranges: Sequence[Tuple[int, int]], events: Sequence,
                          after_bucket=None) -> None:
        """Sum-all-reduce every bucket on the communication stream once its event has fired; the compute
        stream then waits for the communication stream. The 1/world factor is folded into Adam.
        after_bucket(i, stream_ptr): called with the communication stream current right behind bucket i's all-reduce
        (the optimizer step of that bucket: it then overlaps the rest of backward and the later all-reduces)."""
        if not self.enabled:
            return
        cur = torch.cuda.current_stream(self.device)
        cs = self.comm_stream()
        for i, ((off, cnt), ev) in enumerate(zip(ranges, events)):
            cs.wait_event(ev)
            with torch.cuda.stream(cs):
                dist.all_reduce(grads[off:off + cnt], op=dist.ReduceOp.SUM)
                if after_bucket is not None:
                    after_bucket(i, cs.cuda_stream)
        cur.wait_stream(cs)

    def allreduce_flat(self, grads: torch.Tensor, ranges: Sequence[Tuple[int, int]]) -> None:
        """Backend-agnostic version (gloo CPU tests): same buckets, no streams."""
        if not self.enabled:
            return
        for off, cnt in ranges:
            dist.all_reduce(grads[off:off + cnt], op=dist.ReduceOp.SUM)

    def mean(self, t: torch.Tensor) -> torch.Tensor:
        if self.enabled:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t /= self.world
        return t

    def mean_floats(self, vals: Sequence[float]) -> List[float]:
        """Mean over replicas of a few host scalars (epoch logs: every rank must make the SAME callback decisions,
        or a checkpoint's collective on some ranks pairs with a gradient bucket on others)."""
        if not self.enabled:
            return [float(v) for v in vals]
        dev = self.device if dist.get_backend() == 'nccl' else 'cpu'
        t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) / self.world for v in t.cpu().tolist()]

    def max_float(self, v: float) -> float:
        if not self.enabled:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=self.device if torch.cuda.is_available() else 'cpu')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        if self.enabled:
            dist.barrier()
