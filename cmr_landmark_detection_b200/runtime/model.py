"""RvipUNet -- the object create_unet() returns in place of the compiled tf.keras.Model
(src/models/Unets.py:61-133).  It mirrors the slice of the Keras Model API the reference's callers
touch: fit (train_model.py:105-112), predict (predict_model.py:143, predict_4d_on_seg.py:86,
utils/KerasCallbacks.py:481), load_weights / save_weights (predict_model.py:76,
KerasCallbacks.py:54-61), get_weights / set_weights, summary, count_params, optimizer.lr,
stop_training.

PyTorch is used for device memory, streams, pinned host buffers and torch.distributed only; all
arithmetic of the hot path runs in librvip_b200.so through runtime/ffi.py.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ffi
from .dist import DataParallel


def _np_round1(v: float) -> float:
    # Unets.py:105-106 rounds np.float64 values, i.e. numpy's round (x*10, rint, /10)
    return float(np.round(np.float64(v), 1))


def dropout_schedule(config: dict) -> Tuple[List[float], float]:
    depth = config.get('DEPTH', 4)
    d1 = config.get('DROPOUT_MIN', 0.3)
    d3 = config.get('DROPOUT_MAX', 0.5)
    return [_np_round1(v) for v in np.linspace(d1, d3, depth)], float(d3)


class Adam:
    """tf.keras.optimizers.Adam(lr) as get_optimizer builds it (ModelUtils.py:107): beta_1 0.9,
    beta_2 0.999, epsilon 1e-7. `lr` is read and written by ReduceLROnPlateau / LRTensorBoard
    (utils/KerasCallbacks.py:63-70, 173)."""

    def __init__(self, lr: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-7,
                 name: str = 'adam'):
        self.lr = float(lr)
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.iterations = 0
        self.name = name
        self.m = None
        self.v = None

    @property
    def learning_rate(self):
        return self.lr

    @learning_rate.setter
    def learning_rate(self, v):
        self.lr = float(v)

    def get_config(self):
        return {'name': self.name, 'learning_rate': self.lr, 'beta_1': self.beta_1, 'beta_2': self.beta_2,
                'epsilon': self.epsilon}


class SGD:
    """tf.keras.optimizers.SGD as the reference builds it: `SGD(lr, nesterov=True)` for OPTIMIZER='sgd'
    (ModelUtils.py:109-111) and `SGD(name='SGD')` (learning rate 0.01) when OptimizerChanger switches a converged Adam
    run over (utils/KerasCallbacks.py:280-306). Keras update: v = momentum v - lr g; w += nesterov ? momentum v - lr g : v."""

    def __init__(self, lr: float = 0.01, momentum: float = 0.0, nesterov: bool = False, name: str = 'SGD',
                 learning_rate: Optional[float] = None):
        self.lr = float(lr if learning_rate is None else learning_rate)
        self.momentum, self.nesterov = float(momentum), bool(nesterov)
        self.iterations = 0
        self.name = name
        self.velocity = None

    @property
    def learning_rate(self):
        return self.lr

    @learning_rate.setter
    def learning_rate(self, v):
        self.lr = float(v)

    def get_config(self):
        return {'name': self.name, 'learning_rate': self.lr, 'momentum': self.momentum, 'nesterov': self.nesterov}


class History:
    def __init__(self):
        self.history: Dict[str, List[float]] = {}
        self.epoch: List[int] = []


class _Binding:
    """One (batch, training) instantiation of the C handle with its workspace."""

    def __init__(self, model: 'RvipUNet', batch: int, training: bool):
        L = ffi.lib()
        self.batch, self.training = batch, training
        self.h = C.c_void_p()
        ffi.check(L.rvip_create(C.byref(model._cfg), C.byref(self.h)))
        nbytes = L.rvip_workspace_bytes(self.h, batch, int(training))
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=model.device)
        ffi.check(L.rvip_bind(self.h, ffi.ptr(model.params), ffi.ptr(model.grads) if training else None,
                              ffi.ptr(model.bn_state), ffi.ptr(self.workspace), nbytes, batch, int(training)))
        self.packed_version = -1
        self.events: List[torch.cuda.Event] = []
        if training:
            for i in range(L.rvip_num_buckets(self.h)):
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(model.device))   # materialise the cudaEvent_t
                self.events.append(ev)
                ffi.check(L.rvip_set_bucket_event(self.h, i, C.c_void_p(ev.cuda_event)))

    def close(self):
        if self.h:
            ffi.lib().rvip_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RvipUNet:
    def __init__(self, config: dict, name: str = 'unet', device: Optional[torch.device] = None):
        L = ffi.lib()                                   # raises if the CUDA library is not built
        if not torch.cuda.is_available():
            raise ffi.RvipError('RvipUNet needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.name = name
        self.config = dict(config)
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(device)
        dim = config.get('DIM', [224, 224])
        if len(dim) != 2:
            raise NotImplementedError('only 2D DIM is implemented (the RVIP path); got %r' % (dim,))
        for key, want in (('ACTIVATION', 'relu'), ('PAD', 'same')):
            got = config.get(key, want if key == 'PAD' else 'elu')
            if got != want:
                raise NotImplementedError('%s=%r is not implemented (shipped configs use %r)' % (key, got, want))
        if list(config.get('F_SIZE', (3, 3, 3))[-2:]) != [3, 3] or list(config.get('M_POOL', (1, 2, 2))[-2:]) != [2, 2]:
            raise NotImplementedError('only F_SIZE 3x3 and M_POOL 2x2 are implemented')
        drops, dmid = dropout_schedule(config)
        precision = str(config.get('PRECISION', 'bf16')).lower()
        if precision not in ('bf16', 'fp32'):
            raise ValueError('PRECISION must be bf16 or fp32')
        self.precision = precision
        cfg = ffi.rvip_cfg()
        cfg.H, cfg.W = int(dim[0]), int(dim[1])
        cfg.in_ch = int(config.get('IMG_CHANNELS', 1))
        cfg.classes = int(config.get('MASK_CLASSES', 3))
        cfg.depth = int(config.get('DEPTH', 4))
        cfg.filters = int(config.get('FILTERS', 16))
        cfg.batch_norm = int(bool(config.get('BATCH_NORMALISATION', False)))
        cfg.bn_first = int(bool(config.get('BN_FIRST', False)))
        cfg.use_upsample = int(bool(config.get('USE_UPSAMPLE', 'False')))    # string default is truthy (Unets.py:86)
        cfg.precision = 1 if precision == 'bf16' else 0
        for i, d in enumerate(drops):
            cfg.dropout[i] = d
        cfg.dropout_mid = dmid
        cfg.bn_momentum, cfg.bn_eps = 0.99, 1e-3
        self._cfg = cfg
        self.input_shape = (None, cfg.H, cfg.W, cfg.in_ch)
        self.output_shape = (None, cfg.H, cfg.W, cfg.classes)

        # a throw-away handle gives the tensor table (plan only, no device work)
        h = C.c_void_p()
        ffi.check(L.rvip_create(C.byref(cfg), C.byref(h)))
        self.n_params = L.rvip_param_count(h)
        self.n_state = L.rvip_state_count(h)
        self.tensors = []
        for i in range(L.rvip_num_tensors(h)):
            nm = C.create_string_buffer(128)
            is_state, off, nd, dims = C.c_int(), C.c_longlong(), C.c_int(), (C.c_int * 4)()
            ffi.check(L.rvip_tensor_info(h, i, nm, 128, C.byref(is_state), C.byref(off), C.byref(nd), C.byref(dims)))
            self.tensors.append((nm.value.decode(), bool(is_state.value), off.value, tuple(dims[k] for k in range(nd.value))))
        L.rvip_destroy(h)

        self.params = torch.zeros(self.n_params, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(self.n_params, dtype=torch.float32, device=self.device)
        self.bn_state = torch.zeros(self.n_state, dtype=torch.float32, device=self.device)
        self._version = 0
        self._bindings: Dict[Tuple[int, bool], _Binding] = {}
        self._pinned: Dict[Tuple[str, Tuple[int, ...]], torch.Tensor] = {}
        self.optimizer = None          # Adam | SGD
        self.loss_kind = 'mse'
        self.loss_args = {'mask_smaller_than': 0.01}
        self._inplane = None
        self.metrics = []
        self.stop_training = False
        self._step = 0
        self._metric_rows = []
        self._seed = int(config.get('SEED', 42))
        self._loss_dev = torch.zeros(1, dtype=torch.float64, device=self.device)
        # MirroredStrategy equivalent (Unets.py:70-75): on when torch.distributed is initialised with > 1 ranks.
        # DATA_PARALLEL=False builds a rank-local model inside a multi-rank job (no collectives are ever issued).
        self.dp = DataParallel(self.device, enabled=bool(config.get('DATA_PARALLEL', True)))
        self.max_bindings = int(config.get('MAX_BINDINGS', 4))
        # host threads that copy one batch into page-locked memory.  Measured at 8 ranks on one 32-core host
        # (profiles/r2zk_8gpu_staging*.json): 4 threads per rank 52 590 slices/s end to end, 2 threads 50 186, 1 thread 37 906 --
        # the copy (8 x 25 MB per 4.6 ms step) is bound by per-thread memory bandwidth, not by core count
        self._stage_threads = int(os.environ.get('RVIP_STAGING_THREADS') or config.get('STAGING_THREADS', 4))
        self._pool = None
        self._metric_buf = None
        self.set_weights(self._initial_weights(self._seed))

    # ------------------------------------------------------------------ weights
    def _initial_weights(self, seed: int) -> List[np.ndarray]:
        """KERNEL_INIT he_normal for the 3x3 convs (truncated normal, stddev sqrt(2/fan_in)/0.8796),
        glorot_uniform head (Unets.py:128 passes no initializer), zero biases, BN (1, 0, 0, 1)."""
        rng = np.random.default_rng(seed)
        out = []
        for name, is_state, off, shape in self.tensors:
            leaf = name.rsplit('/', 1)[-1]
            if leaf == 'kernel':
                kh, kw, cin, cout = shape
                if name.startswith('head/'):
                    lim = math.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
                    w = rng.uniform(-lim, lim, size=shape)
                else:
                    std = math.sqrt(2.0 / (kh * kw * cin)) / 0.87962566103423978
                    w = rng.standard_normal(size=shape)
                    bad = np.abs(w) > 2.0
                    while bad.any():
                        w[bad] = rng.standard_normal(size=int(bad.sum()))
                        bad = np.abs(w) > 2.0
                    w *= std
                out.append(w.astype(np.float32))
            elif leaf in ('gamma', 'moving_variance'):
                out.append(np.ones(shape, np.float32))
            else:
                out.append(np.zeros(shape, np.float32))
        return out

    def get_weights(self) -> List[np.ndarray]:
        p = self.params.detach().cpu().numpy()
        s = self._synced_state().cpu().numpy()
        out = []
        for name, is_state, off, shape in self.tensors:
            n = int(np.prod(shape))
            out.append((s if is_state else p)[off:off + n].reshape(shape).copy())
        return out

    def set_weights(self, weights: Sequence[np.ndarray]):
        if len(weights) != len(self.tensors):
            raise ValueError('expected %d weight tensors, got %d' % (len(self.tensors), len(weights)))
        p = np.zeros(self.n_params, np.float32)
        s = np.zeros(self.n_state, np.float32)
        for (name, is_state, off, shape), w in zip(self.tensors, weights):
            w = np.asarray(w, dtype=np.float32)
            if tuple(w.shape) != tuple(shape):
                raise ValueError('weight %s: expected shape %s, got %s' % (name, shape, w.shape))
            (s if is_state else p)[off:off + w.size] = w.ravel()
        self.params.copy_(torch.from_numpy(p))
        self.bn_state.copy_(torch.from_numpy(s))
        self._version += 1

    def _synced_state(self) -> torch.Tensor:
        """MirroredStrategy keeps BN moving statistics per replica and averages them on read
        (synchronization=ON_READ, aggregation=MEAN)."""
        return self.dp.mean(self.bn_state.clone()) if self.dp.world > 1 else self.bn_state

    def keras_layer_names(self) -> List[Tuple[str, List[int]]]:
        """[(keras layer name, indices into self.tensors)] in creation order."""
        out, n_conv, n_bn, i = [], 0, 0, 0
        while i < len(self.tensors):
            name = self.tensors[i][0]
            if name.endswith('/kernel'):
                if name.startswith('head/'):
                    lname = self.name if self.name else 'unet'
                    lname = 'unet'
                else:
                    lname = 'conv2d' if n_conv == 0 else 'conv2d_%d' % n_conv
                    n_conv += 1
                out.append((lname, [i, i + 1]))
                i += 2
            else:
                lname = 'batch_normalization' if n_bn == 0 else 'batch_normalization_%d' % n_bn
                n_bn += 1
                out.append((lname, [i, i + 1, i + 2, i + 3]))
                i += 4
        return out

    _LEAF = {'kernel': 'kernel:0', 'bias': 'bias:0', 'gamma': 'gamma:0', 'beta': 'beta:0',
             'moving_mean': 'moving_mean:0', 'moving_variance': 'moving_variance:0'}

    def save_weights(self, filepath: str, overwrite: bool = True):
        """model.save_weights (ModelCheckpoint(save_weights_only=True), KerasCallbacks.py:54-61).  A path ending in .h5 /
        .hdf5 is written as a Keras HDF5 weight file (tf.keras hdf5_format layout: root attributes layer_names / backend /
        keras_version, one group per layer with weight_names and <layer>/<var>:0 datasets) by utils/hdf5_lite.py -- h5py is
        not needed.  Any other path is written as <path>.npz with the same keys."""
        ws = self.get_weights()               # a collective under data parallelism (moving statistics): every rank calls it
        layers = []
        for lname, idxs in self.keras_layer_names():
            layers.append((lname, [('%s/%s' % (lname, self._LEAF[self.tensors[i][0].rsplit('/', 1)[-1]]), ws[i]) for i in idxs]))
        if self.dp.rank != 0:
            return
        os.makedirs(os.path.dirname(os.path.abspath(filepath)), exist_ok=True)
        if filepath.endswith(('.h5', '.hdf5')):
            from ..utils import hdf5_lite
            hdf5_lite.save_keras_weights(filepath, layers)
            return
        blob = {'%s/%s' % (lname, wn): arr for lname, weights in layers for wn, arr in weights}
        blob['layer_names'] = np.array([lname for lname, _ in layers])
        np.savez(filepath if filepath.endswith('.npz') else filepath + '.npz', **blob)

    def load_weights(self, filepath: str):
        """model.load_weights (predict_model.py:76): a Keras HDF5 weight file (read by utils/hdf5_lite.py; files written
        with h5py's default settings, no compression) or the .npz container save_weights writes.  Like Keras' topological
        loading, layers are matched by POSITION among the layers that have weights (auto-numbered layer names depend on
        the process history) and every shape is checked."""
        from ..utils import hdf5_lite
        path = filepath
        if not os.path.exists(path) and os.path.exists(filepath + '.npz'):
            path = filepath + '.npz'
        with open(path, 'rb') as fh:
            head = fh.read(8)
        mine = self.keras_layer_names()
        if head == hdf5_lite.SIGNATURE or path.endswith(('.h5', '.hdf5')):
            file_layers = hdf5_lite.load_keras_weights(path)
        else:
            z = np.load(path, allow_pickle=False)
            file_layers = []
            for fl in [str(s) for s in z['layer_names']]:
                prefix = fl + '/'
                # this layer's arrays, in the order the model lists its own (kernel, bias | gamma, beta, moving_*)
                file_layers.append((fl, [(k[len(prefix):], z[k]) for k in z.files if k.startswith(prefix)]))
        if len(file_layers) != len(mine):
            raise ValueError('weight file has %d layers with weights, model has %d' % (len(file_layers), len(mine)))
        ws = [None] * len(self.tensors)
        for (fl, weights), (lname, idxs) in zip(file_layers, mine):
            by_leaf = {wn.rsplit('/', 1)[-1]: arr for wn, arr in weights}
            if len(weights) != len(idxs):
                raise ValueError('layer %s: file has %d weight arrays, model layer %s has %d' % (fl, len(weights), lname, len(idxs)))
            for i in idxs:
                leaf = self._LEAF[self.tensors[i][0].rsplit('/', 1)[-1]]
                if leaf not in by_leaf:
                    raise ValueError('layer %s in the weight file has no %s' % (fl, leaf))
                arr = np.asarray(by_leaf[leaf], dtype=np.float32)
                if tuple(arr.shape) != tuple(self.tensors[i][3]):
                    raise ValueError('%s/%s: shape %s in the file, %s in the model' % (fl, leaf, arr.shape, tuple(self.tensors[i][3])))
                ws[i] = arr
        self.set_weights(ws)

    def count_params(self) -> int:
        return int(self.n_params + self.n_state)

    def summary(self, print_fn=None, line_length=None):
        print_fn = print_fn or print
        print_fn('Model: "%s"  (B200-native, precision=%s)' % (self.name, self.precision))
        print_fn('%-28s %-22s %12s' % ('Layer', 'Weights shape', 'Param #'))
        for lname, idxs in self.keras_layer_names():
            n = sum(int(np.prod(self.tensors[i][3])) for i in idxs)
            print_fn('%-28s %-22s %12d' % (lname, str(self.tensors[idxs[0]][3]), n))
        print_fn('Total params: {:,}'.format(self.count_params()))
        print_fn('Trainable params: {:,}'.format(self.n_params))
        print_fn('Non-trainable params: {:,}'.format(self.n_state))

    # ------------------------------------------------------------------ compile
    def compile(self, optimizer=None, loss=None, metrics=None, **kw):
        if optimizer is not None:
            if not isinstance(optimizer, (Adam, SGD)):
                raise NotImplementedError('only the Adam and SGD optimizers are implemented on the device path')
            self.optimizer = optimizer
        from ..models.Loss_and_metrics import resolve_loss
        loss = resolve_loss(loss)
        if loss is not None:
            self.loss_kind = loss.rvip_kind
            self.loss_args = dict(getattr(loss, 'rvip_args', None) or {'mask_smaller_than': 0.01})
        self.metrics = []
        for m in (metrics or []):
            if not hasattr(m, 'rvip_channels'):
                raise NotImplementedError('metric %r is not implemented on the device path (the dice_coef* family of '
                                          'models/Loss_and_metrics.py is)' % (m,))
            m.rvip_channels(self._cfg.classes)       # raises for a channel the model does not have
            self.metrics.append(m)
        if self.loss_kind == 'weighted':
            H, W = self._cfg.H, self._cfg.W
            yy, xx = np.mgrid[0:H, 0:W]
            d = np.minimum(np.minimum(yy, xx), np.minimum(H - 1 - yy, W - 1 - xx))
            ramp = np.linspace(0, 100, H // 2)                      # Loss_and_metrics.py:62-69
            w = ramp[np.minimum(d, H // 2 - 1)].astype(np.float32)
            w[d == 0] = 0.0
            self._inplane = torch.from_numpy(w).to(self.device)
        return self

    # ------------------------------------------------------------------ plumbing
    def _binding(self, batch: int, training: bool) -> _Binding:
        key = (batch, training)
        b = self._bindings.pop(key, None)
        if b is None:
            # every binding owns a full workspace (GBs at training sizes): keep the most recently used few
            # (ragged last batches, volumes of varying depth) and release the rest
            while len(self._bindings) >= max(self.max_bindings, 1):
                old_key = next(iter(self._bindings))
                torch.cuda.current_stream(self.device).synchronize()
                self._bindings.pop(old_key).close()
            b = _Binding(self, batch, training)
        self._bindings[key] = b           # dict order = least recently used first
        if b.packed_version != self._version:
            ffi.check(ffi.lib().rvip_pack_weights(b.h, self._stream()))
            b.packed_version = self._version
        return b

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _pin(self, tag: str, shape, dtype=torch.float32) -> torch.Tensor:
        key = (tag, tuple(shape))
        t = self._pinned.get(key)
        if t is None:
            t = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
            self._pinned[key] = t
        return t

    def _stage(self, dst: torch.Tensor, src: np.ndarray) -> torch.Tensor:
        """Host batch -> page-locked memory for the asynchronous H2D copy.  Already page-locked sources (cudaHostRegister /
        pinned torch storage behind the ndarray) are used in place; anything else is copied into the pinned slot `dst` by a
        few threads (numpy releases the GIL): one thread moves ~5 GB/s, i.e. 25 MB per step cost as much as the GPU step
        itself once eight ranks share a host (round-1 finding: N = 8 end-to-end 12 % under the device-resident rate)."""
        t = torch.from_numpy(src)
        try:
            if t.is_pinned():
                return t
        except RuntimeError:
            pass
        n = src.shape[0]
        d = dst.numpy()
        workers = min(self._stage_threads, n)
        if workers <= 1 or src.nbytes < (1 << 20):
            np.copyto(d, src)
            return dst
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=self._stage_threads)
        bounds = [n * i // workers for i in range(workers + 1)]
        futs = [self._pool.submit(np.copyto, d[bounds[i]:bounds[i + 1]], src[bounds[i]:bounds[i + 1]]) for i in range(workers)]
        for f in futs:
            f.result()
        return dst

    def _check_x(self, x):
        if x.ndim != 4 or tuple(x.shape[1:]) != (self._cfg.H, self._cfg.W, self._cfg.in_ch):
            raise ValueError('expected input [N,%d,%d,%d], got %s' % (self._cfg.H, self._cfg.W, self._cfg.in_ch,
                                                                        tuple(x.shape)))

    def _check_y(self, y, n: int):
        if y.ndim != 4 or tuple(y.shape) != (n, self._cfg.H, self._cfg.W, self._cfg.classes):
            raise ValueError('expected target [%d,%d,%d,%d], got %s' % (n, self._cfg.H, self._cfg.W, self._cfg.classes,
                                                                         tuple(y.shape)))

    def _check_dev(self, t: torch.Tensor, what: str):
        # raw pointers cross the C ABI: the kernels read dense fp32 NHWC on this model's device
        if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
            raise ValueError('%s must be a contiguous float32 tensor on %s (got %s, %s, contiguous=%s)'
                             % (what, self.device, t.dtype, t.device, t.is_contiguous()))

    # ------------------------------------------------------------------ inference
    def predict_device(self, x_dev: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x_dev [B,H,W,in_ch] fp32 on the device -> heat [B,H,W,classes] fp32 (device)."""
        B = x_dev.shape[0]
        self._check_x(x_dev)
        self._check_dev(x_dev, 'x')
        b = self._binding(B, False)
        if out is None:
            out = torch.empty((B, self._cfg.H, self._cfg.W, self._cfg.classes), dtype=torch.float32, device=self.device)
        ffi.check(ffi.lib().rvip_predict(b.h, ffi.ptr(x_dev), ffi.ptr(out), self._stream()))
        return out

    def predict(self, x, batch_size: Optional[int] = None, verbose=0, steps=None, **kw) -> np.ndarray:
        """model.predict: ndarray [N,H,W,C] (default batch 32) or a Sequence whose items are x or (x, y).
        Host side: batch i + 1 is staged in page-locked memory and copied to the device on a copy stream while batch i
        runs; the heat maps go device -> ONE page-locked result buffer (async, no per-batch synchronisation) whose numpy
        view is returned, so the only host copy is the staging of the inputs."""
        if isinstance(x, np.ndarray):
            self._check_x(x)
            bs = int(batch_size or 32)
            batches = [x[i:i + bs] for i in range(0, x.shape[0], bs)]
        else:
            n = len(x) if steps is None else steps
            batches = [(x[i][0] if isinstance(x[i], (tuple, list)) else x[i]) for i in range(n)]
        batches = [np.ascontiguousarray(xb, dtype=np.float32) for xb in batches]
        total = sum(len(xb) for xb in batches)
        shape = (total, self._cfg.H, self._cfg.W, self._cfg.classes)
        if total == 0:
            return np.zeros(shape, np.float32)
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            if not hasattr(self, '_pcopy_stream'):
                self._pcopy_stream = torch.cuda.Stream(device=self.device)
            cs = self._pcopy_stream
            out = torch.empty(shape, dtype=torch.float32, pin_memory=True)     # torch caches page-locked blocks
            ready = [torch.cuda.Event() for _ in range(2)]
            used = [torch.cuda.Event() for _ in range(2)]
            dev_in: Dict[Tuple[int, Tuple[int, ...]], torch.Tensor] = {}
            pos = 0
            for i, xb in enumerate(batches):
                self._check_x(xb)
                s = i & 1
                if i >= 2:
                    ready[s].synchronize()                  # the pinned slot's previous H2D has left it
                hp = self._stage(self._pin('px%d' % s, xb.shape), xb)
                key = (s, tuple(xb.shape))
                if key not in dev_in:
                    dev_in[key] = torch.empty(xb.shape, dtype=torch.float32, device=self.device)
                xd = dev_in[key]
                if i >= 2:
                    cs.wait_event(used[s])                  # the forward that last read this device slot has finished
                with torch.cuda.stream(cs):
                    xd.copy_(hp, non_blocking=True)
                    ready[s].record(cs)
                main.wait_event(ready[s])
                heat = self.predict_device(xd)
                used[s].record(main)
                out[pos:pos + len(xb)].copy_(heat, non_blocking=True)
                pos += len(xb)
            main.synchronize()
        return out.numpy()

    def __call__(self, x, training=False):
        return self.predict(np.asarray(x))

    # ------------------------------------------------------------------ training
    def train_step_device(self, x_dev: torch.Tensor, y_dev: torch.Tensor, heat: Optional[torch.Tensor] = None,
                          apply_optimizer: bool = True) -> torch.Tensor:
        """One optimisation step on device-resident data. Returns the (device, float64) mean loss."""
        if self.optimizer is None:
            raise RuntimeError('compile(optimizer=...) first')
        L = ffi.lib()
        B = x_dev.shape[0]
        self._check_x(x_dev)
        self._check_y(y_dev, B)
        self._check_dev(x_dev, 'x')
        self._check_dev(y_dev, 'y')
        b = self._binding(B, True)
        if heat is None:
            heat = self._heat_buf(B)
        self._last_heat = heat
        self._step += 1
        seed = (self._seed * 1000003 + self._step) ^ (self.dp.rank << 40)
        thr = float(self.loss_args.get('mask_smaller_than', 0.01))
        if self.loss_kind == 'bce_dice':
            ffi.check(L.rvip_set_loss_weights(b.h, float(self.loss_args.get('w_bce', 1.0)),
                                              float(self.loss_args.get('w_dice', 1.0))))
        # single replica + Adam: the step applies the optimizer itself, bucket by bucket behind the backward pass
        bucketed = (apply_optimizer and isinstance(self.optimizer, Adam) and not os.environ.get('RVIP_NO_INLINE_ADAM'))
        inline = bucketed and self.dp.world == 1
        if bucketed and self.optimizer.m is None:
            # the moments are zero-filled on THIS stream before the step is queued: the per-bucket optimizer work runs on
            # other streams that are only ordered behind the step's own events
            self.optimizer.m = torch.zeros_like(self.params)
            self.optimizer.v = torch.zeros_like(self.params)
        if inline:
            opt = self.optimizer
            opt.iterations += 1
            ffi.check(L.rvip_set_inline_adam(b.h, ffi.ptr(opt.m), ffi.ptr(opt.v), opt.lr, opt.beta_1, opt.beta_2,
                                             opt.epsilon, opt.iterations, 1.0))
        ffi.check(L.rvip_train_step(b.h, ffi.ptr(x_dev), ffi.ptr(y_dev), ffi.ptr(self._inplane),
                                    ffi.LOSS_KINDS[self.loss_kind], thr, C.c_uint64(seed & (2 ** 64 - 1)),
                                    ffi.ptr(heat), ffi.ptr(self._loss_dev), self._stream()))
        if inline:
            self._version += 1
            b.packed_version = self._version      # the step re-packed this binding's operand copies
            return self._loss_dev
        if self.dp.world > 1:
            if bucketed:
                # data parallel: every bucket is stepped on the communication stream right behind its all-reduce
                opt = self.optimizer
                opt.iterations += 1

                def step_bucket(i, stream_ptr):
                    ffi.check(L.rvip_adam_bucket(b.h, i, ffi.ptr(opt.m), ffi.ptr(opt.v), opt.lr, opt.beta_1, opt.beta_2,
                                                 opt.epsilon, opt.iterations, 1.0 / self.dp.world, C.c_void_p(stream_ptr)))
                self.dp.allreduce_buckets(self.grads, self._buckets(b), b.events, after_bucket=step_bucket)
                self._version += 1
                b.packed_version = self._version
                return self._loss_dev
            self.dp.allreduce_buckets(self.grads, self._buckets(b), b.events)
        if apply_optimizer:
            self.apply_gradients(b)
        return self._loss_dev

    def apply_gradients(self, b: Optional[_Binding] = None):
        opt = self.optimizer
        if b is None:
            b = next(v for k, v in self._bindings.items() if k[1])
        opt.iterations += 1
        if isinstance(opt, SGD):
            if opt.momentum != 0.0 and opt.velocity is None:
                opt.velocity = torch.zeros_like(self.params)
            ffi.check(ffi.lib().rvip_sgd_step(b.h, ffi.ptr(opt.velocity), opt.lr, opt.momentum, int(opt.nesterov),
                                              1.0 / self.dp.world, self._stream()))
        else:
            if opt.m is None:
                opt.m = torch.zeros_like(self.params)
                opt.v = torch.zeros_like(self.params)
            ffi.check(ffi.lib().rvip_adam_step(b.h, ffi.ptr(opt.m), ffi.ptr(opt.v), opt.lr, opt.beta_1, opt.beta_2,
                                               opt.epsilon, opt.iterations, 1.0 / self.dp.world, self._stream()))
        self._version += 1
        b.packed_version = self._version      # rvip_adam_step re-packs this binding's operand copies

    def _buckets(self, b: _Binding):
        if not hasattr(b, 'bucket_ranges'):
            L = ffi.lib()
            r = []
            for i in range(L.rvip_num_buckets(b.h)):
                o, c = C.c_longlong(), C.c_longlong()
                ffi.check(L.rvip_bucket(b.h, i, C.byref(o), C.byref(c)))
                r.append((o.value, c.value))
            b.bucket_ranges = r
        return b.bucket_ranges

    def _heat_buf(self, B):
        key = ('heat', B)
        if key not in self._pinned:
            self._pinned[key] = torch.empty((B, self._cfg.H, self._cfg.W, self._cfg.classes), dtype=torch.float32,
                                            device=self.device)
        return self._pinned[key]

    def train_on_batch(self, x: np.ndarray, y: np.ndarray) -> float:
        """Host-facing step: pinned staging + H2D of (x, y), device step, D2H of the loss."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        self._check_x(x)
        self._check_y(y, len(x))
        with torch.cuda.device(self.device):
            hx = self._pin('tx', x.shape)
            hy = self._pin('ty', y.shape)
            hx.copy_(torch.from_numpy(x))
            hy.copy_(torch.from_numpy(y))
            xd = hx.to(self.device, non_blocking=True)
            yd = hy.to(self.device, non_blocking=True)
            loss = self.train_step_device(xd, yd)
            return float(loss.item())

    def _run_steps(self, batches, with_metrics: bool = False) -> List[float]:
        """Pipelined training steps over an iterable of host (x, y) batches -- the loop inside fit().

        Three pinned staging slots and a staging thread: while the GPU runs step i and this thread queues the launches of
        step i + 1, the staging thread pulls batch i + 2 from the generator and copies it into page-locked memory; a copy
        stream moves a staged batch to the device behind the running step; the loss of step i is read back (async D2H +
        event) only after step i + 1 has been queued.  Every step still pays its own H2D copy of (x, y) and its own D2H
        read of the loss; the device never waits for the host unless generator + staging alone exceed a step.
        `self.last_fit_timing` holds where this thread's time went (ms per step): waiting for a staged batch, queueing
        the step's launches, waiting for the previous step's loss."""
        import queue
        import threading
        losses: List[float] = []
        self._metric_rows: List[int] = []       # pixels per step whose heat statistics sit in self._metric_buf
        nrow = 1 + 3 * self._cfg.classes
        NS = 3
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            if not hasattr(self, '_copy_stream'):
                self._copy_stream = torch.cuda.Stream(device=self.device)
                self._slot_ev = [dict(ready=torch.cuda.Event(), used=torch.cuda.Event()) for _ in range(NS)]
                self._loss_ev = [torch.cuda.Event() for _ in range(2)]
                self._loss_pin = [torch.zeros(1, dtype=torch.float64, pin_memory=True) for _ in range(2)]
            cs = self._copy_stream
            staged: 'queue.Queue' = queue.Queue(maxsize=NS - 1)
            free: 'queue.Queue' = queue.Queue()
            for s in range(NS):
                free.put((s, False))
            stop = threading.Event()

            def producer():
                try:
                    for x, y in batches:
                        x = np.ascontiguousarray(x, dtype=np.float32)
                        y = np.ascontiguousarray(y, dtype=np.float32)
                        self._check_x(x)
                        self._check_y(y, len(x))
                        s, used = free.get()
                        if stop.is_set():
                            return
                        if used:
                            self._slot_ev[s]['ready'].synchronize()   # the slot's previous H2D has left pinned memory
                        hx = self._stage(self._pin('fx%d' % s, x.shape), x)
                        hy = self._stage(self._pin('fy%d' % s, y.shape), y)
                        staged.put((s, hx, hy))
                    staged.put(None)
                except BaseException as e:          # surface generator / validation errors in the training thread
                    staged.put(e)

            th = threading.Thread(target=producer, daemon=True)
            th.start()
            dev_slots: Dict[Tuple[int, Tuple[int, ...], Tuple[int, ...]], Tuple[torch.Tensor, torch.Tensor]] = {}
            used_once = [False] * NS
            pending = None       # loss slot whose value has not been read yet
            t_wait = t_launch = t_loss = 0.0
            i = 0
            try:
                while True:
                    t0 = time.perf_counter()
                    item = staged.get()
                    t1 = time.perf_counter()
                    if item is None:
                        break
                    if isinstance(item, BaseException):
                        raise item
                    s, hx, hy = item
                    ev = self._slot_ev[s]
                    key = (s, tuple(hx.shape), tuple(hy.shape))
                    if key not in dev_slots:
                        dev_slots[key] = (torch.empty(tuple(hx.shape), dtype=torch.float32, device=self.device),
                                          torch.empty(tuple(hy.shape), dtype=torch.float32, device=self.device))
                    xd, yd = dev_slots[key]
                    if used_once[s]:
                        cs.wait_event(ev['used'])       # the step that last read this device slot has finished
                    with torch.cuda.stream(cs):
                        xd.copy_(hx, non_blocking=True)
                        yd.copy_(hy, non_blocking=True)
                        ev['ready'].record(cs)
                    free.put((s, True))                 # the staging thread may refill the slot once `ready` has fired
                    main.wait_event(ev['ready'])
                    loss_dev = self.train_step_device(xd, yd)
                    if with_metrics:
                        # training metrics: one reduction pass over this step's heat map and target, rows stay on the device
                        if self._metric_buf is None or self._metric_buf.shape[0] <= i:
                            grown = torch.zeros((max(256, 2 * (i + 1)), nrow), dtype=torch.float64, device=self.device)
                            if self._metric_buf is not None:
                                grown[:self._metric_buf.shape[0]] = self._metric_buf
                            self._metric_buf = grown
                        self._stats_row(self._last_heat, yd, self._metric_buf[i])
                        self._metric_rows.append(int(hx.shape[0]) * self._cfg.H * self._cfg.W)
                    ev['used'].record(main)
                    ls = i & 1
                    self._loss_pin[ls].copy_(loss_dev.reshape(1), non_blocking=True)
                    self._loss_ev[ls].record(main)
                    used_once[s] = True
                    t2 = time.perf_counter()
                    if pending is not None:
                        self._loss_ev[pending].synchronize()
                        losses.append(float(self._loss_pin[pending][0]))
                    pending = ls
                    t3 = time.perf_counter()
                    t_wait += t1 - t0
                    t_launch += t2 - t1
                    t_loss += t3 - t2
                    i += 1
                if pending is not None:
                    self._loss_ev[pending].synchronize()
                    losses.append(float(self._loss_pin[pending][0]))
            finally:
                stop.set()
                free.put((0, False))       # unblock a producer waiting for a slot
                th.join(timeout=5.0)
        n = max(i, 1)
        self.last_fit_timing = {'steps': i, 'wait_staged_ms': round(t_wait / n * 1e3, 3),
                                'queue_launches_ms': round(t_launch / n * 1e3, 3),
                                'wait_prev_loss_ms': round(t_loss / n * 1e3, 3)}
        return losses

    # ------------------------------------------------------------------ loss / metric bookkeeping
    def _metric_names(self) -> List[str]:
        return [m.__name__ for m in self.metrics]

    def _stats_row(self, heat: torch.Tensor, y_dev: torch.Tensor, out: torch.Tensor):
        """rvip_heat_stats: out[0] = sum of the compiled loss's per-pixel terms, out[1 + 3c ..] = {sum t p, sum p, sum t}."""
        thr = float(self.loss_args.get('mask_smaller_than', 0.01))
        n_pix = heat.shape[0] * self._cfg.H * self._cfg.W
        ffi.check(ffi.lib().rvip_heat_stats(ffi.ptr(heat), ffi.ptr(y_dev), ffi.ptr(self._inplane), n_pix,
                                            self._cfg.H * self._cfg.W, self._cfg.classes,
                                            ffi.LOSS_KINDS[self.loss_kind], thr, ffi.ptr(out), self._stream()))

    def _finish_stats(self, row: np.ndarray, n_pix: int) -> Dict[str, float]:
        """Host end of rvip_heat_stats for ONE batch: the loss value Keras logs and every compiled metric."""
        C_ = self._cfg.classes
        tp, sp, st = row[1::3][:C_], row[2::3][:C_], row[3::3][:C_]
        out = {'loss': float(row[0]) / n_pix}
        if self.loss_kind == 'bce_dice':
            dice = (2.0 * tp.sum() + 1.0) / (st.sum() + sp.sum() + 1.0)
            out['loss'] = float(self.loss_args.get('w_bce', 1.0)) * out['loss'] - float(self.loss_args.get('w_dice', 1.0)) * dice
        for m in self.metrics:
            ch = m.rvip_channels(C_)
            out[m.__name__] = float((2.0 * tp[ch].sum() + 1.0) / (st[ch].sum() + sp[ch].sum() + 1.0))
        return out

    def _shard(self, x, y=None, per_replica: bool = False):
        """MirroredStrategy shards every global batch over the replicas (Unets.py:70-75): rank r owns samples
        [r * per, (r + 1) * per).  per_replica: the caller already hands out this rank's own batches."""
        if self.dp.world == 1 or per_replica:
            return (x, y)
        from .dist import shard_range
        lo, hi = shard_range(len(x), self.dp.rank, self.dp.world)
        return (x[lo:hi], None if y is None else y[lo:hi])

    def evaluate(self, x, y=None, batch_size=32, verbose=0, return_dict=False):
        """Validation loss (+ compiled metrics) in inference mode, Keras' batch-size-weighted mean over batches: device
        forward (rvip_predict), then ONE reduction kernel per batch over heat map and target (rvip_heat_stats); only the
        1 + 3 C sums come back to the host."""
        if isinstance(x, np.ndarray):
            items = ((x[i:i + batch_size], y[i:i + batch_size]) for i in range(0, len(x), batch_size))
            per_replica = False
        else:
            items = (x[i][:2] for i in range(len(x)))
            per_replica = bool(getattr(x, 'rvip_per_replica', False))
        tot: Dict[str, float] = {}
        n = 0
        nrow = 1 + 3 * self._cfg.classes
        with torch.cuda.device(self.device):
            row_dev = torch.zeros(nrow, dtype=torch.float64, device=self.device)
            for xb, yb in items:
                xb, yb = self._shard(xb, yb, per_replica)
                if len(xb) == 0:
                    continue
                xb = np.ascontiguousarray(xb, dtype=np.float32)
                yb = np.ascontiguousarray(yb, dtype=np.float32)
                self._check_x(xb)
                self._check_y(yb, len(xb))
                heat = self.predict_device(torch.from_numpy(xb).to(self.device))
                self._stats_row(heat, torch.from_numpy(yb).to(self.device), row_dev)
                vals = self._finish_stats(row_dev.cpu().numpy(), len(xb) * self._cfg.H * self._cfg.W)
                for k, v in vals.items():
                    tot[k] = tot.get(k, 0.0) + v * len(xb)
                n += len(xb)
        keys = ['loss'] + self._metric_names()
        vals = self.dp.mean_floats([tot.get(k, 0.0) for k in keys] + [float(n)])
        res = {k: v / max(vals[-1], 1e-30) for k, v in zip(keys, vals[:-1])}
        return res if return_dict else res['loss']

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, callbacks=None, validation_data=None,
            shuffle=True, initial_epoch=0, steps_per_epoch=None, max_queue_size=10, workers=1,
            use_multiprocessing=False, **kw) -> History:
        """Epoch/step loop with the Keras callback protocol (train_model.py:105-112). `x` is an ndarray
        (with `y`) or a keras.utils.Sequence-like object (len / getitem -> (x, y) / on_epoch_end).
        Data parallel (one process per GPU): every rank iterates the same global batches and trains on its shard;
        the epoch logs are averaged over the ranks before any callback sees them, so checkpoint / learning-rate /
        early-stopping decisions (and the collectives behind save_weights) are identical everywhere."""
        hist = History()
        callbacks = list(callbacks or [])
        for cb in callbacks:
            if hasattr(cb, 'set_model'):
                cb.set_model(self)
        self.stop_training = False
        for cb in callbacks:
            if hasattr(cb, 'on_train_begin'):
                cb.on_train_begin({})
        is_seq = not isinstance(x, np.ndarray)
        per_replica = bool(getattr(x, 'rvip_per_replica', False)) if is_seq else False
        bs = int(batch_size or 32)
        rng = np.random.default_rng(self._seed)        # same permutation on every rank: identical global batches
        mnames = self._metric_names()
        for epoch in range(initial_epoch, epochs):
            for cb in callbacks:
                if hasattr(cb, 'on_epoch_begin'):
                    cb.on_epoch_begin(epoch, {})
            t0 = time.time()
            if is_seq:
                n_steps = len(x) if steps_per_epoch is None else steps_per_epoch
                gen = (self._shard(*tuple(x[i][:2]), per_replica) for i in range(n_steps))
            else:
                order = rng.permutation(len(x)) if shuffle else np.arange(len(x))
                n_steps = -(-len(x) // bs) if steps_per_epoch is None else steps_per_epoch     # Keras: ceil
                gen = (self._shard(x[order[i * bs:(i + 1) * bs]], y[order[i * bs:(i + 1) * bs]])
                       for i in range(max(n_steps, 1)))
            losses = self._run_steps(gen, with_metrics=bool(mnames))
            logs = {'loss': float(np.mean(losses)) if losses else float('nan')}
            for k, v in self._epoch_metrics().items():
                logs[k] = v
            if validation_data is not None:
                if isinstance(validation_data, (tuple, list)):
                    val = self.evaluate(validation_data[0], validation_data[1], batch_size=bs, return_dict=True)
                else:
                    val = self.evaluate(validation_data, return_dict=True)
                for k, v in val.items():
                    logs['val_' + k] = v
            if self.dp.world > 1:
                train_keys = [k for k in logs if not k.startswith('val_')]      # val_* are already global
                for k, v in zip(train_keys, self.dp.mean_floats([logs[k] for k in train_keys])):
                    logs[k] = v
            logs['lr'] = self.optimizer.lr
            if is_seq and hasattr(x, 'on_epoch_end'):
                x.on_epoch_end()
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(v)
            if verbose and self.dp.rank == 0:
                print('Epoch %d/%d - %.1fs - %s' % (epoch + 1, epochs, time.time() - t0,
                                                    ' - '.join('%s: %.6g' % kv for kv in logs.items())))
            for cb in callbacks:
                if hasattr(cb, 'on_epoch_end'):
                    cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        for cb in callbacks:
            if hasattr(cb, 'on_train_end'):
                cb.on_train_end({})
        self.history = hist
        return hist

    def _epoch_metrics(self) -> Dict[str, float]:
        """Keras logs a compiled metric as the mean of its per-batch values: finish the rows rvip_heat_stats left on the
        device during the epoch's training steps."""
        if not self.metrics or not self._metric_rows:
            return {}
        rows = self._metric_buf[:len(self._metric_rows)].cpu().numpy()
        acc: Dict[str, float] = {}
        for row, n_pix in zip(rows, self._metric_rows):
            for k, v in self._finish_stats(row, n_pix).items():
                if k != 'loss':
                    acc[k] = acc.get(k, 0.0) + v
        return {k: v / len(rows) for k, v in acc.items()}

    # ------------------------------------------------------------------ introspection (tests / bench)
    def debug_buffer(self, layer: str, which: int, batch: int, training: bool) -> Optional[torch.Tensor]:
        b = self._bindings[(batch, training)]
        p, n, es = C.c_void_p(), C.c_longlong(), C.c_int()
        ffi.check(ffi.lib().rvip_debug_buffer(b.h, layer.encode(), which, C.byref(p), C.byref(n), C.byref(es)))
        if not p.value or n.value == 0:
            return None
        dt = torch.bfloat16 if es.value == 2 else torch.float32
        off = p.value - b.workspace.data_ptr()          # every buffer lives inside the bound workspace
        torch.cuda.current_stream(self.device).synchronize()
        return b.workspace[off:off + n.value * es.value].view(dt).clone()

    def profile(self, batch: int, training: bool, enable: bool):
        ffi.check(ffi.lib().rvip_profile(self._bindings[(batch, training)].h, int(enable)))

    def profile_read(self, batch: int, training: bool) -> Dict[str, Tuple[float, int]]:
        L = ffi.lib()
        ms = (C.c_float * ffi.NUM_KERNEL_CLASSES)()
        n = (C.c_longlong * ffi.NUM_KERNEL_CLASSES)()
        ffi.check(L.rvip_profile_read(self._bindings[(batch, training)].h, C.byref(ms), C.byref(n)))
        return {L.rvip_kernel_class_name(i).decode(): (float(ms[i]), int(n[i])) for i in range(ffi.NUM_KERNEL_CLASSES)}

    def profile_detail(self, batch: int, training: bool):
        """Per launch group (class, 'layer:op', ms) of the log consumed by the last profile_read()."""
        txt = ffi.lib().rvip_profile_detail(self._bindings[(batch, training)].h).decode()
        rows = [ln.split(',') for ln in txt.splitlines() if ln]
        return [(c, tag, float(ms)) for c, tag, ms in rows]

    def launch_count(self) -> int:
        return int(sum(ffi.lib().rvip_launch_count(b.h) for b in self._bindings.values()))
