"""CPU ORACLE for the U-Net half of the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (cmr_landmark_detection_b200) never
does; it fails loudly when its CUDA library is missing.

PARITY UNPINNED (numerically): the reference computes this path inside
tensorflow==2.3.0 (environment.yml:126) which is not importable here and whose source
is not under /root/reference.  This file restates the reference's graph from its own
call sites and the TF-2.3 layer semantics listed in SURVEY.md Appendix C:

  topology ............ src/models/Unets.py:755-869 (unet), :61-133 (create_unet, head :128)
  conv block .......... src/models/KerasLayers.py:660-693 (conv_layer_fn; BN_FIRST False
                        => Conv+bias -> ReLU -> BatchNorm, True => Conv -> BN -> ReLU)
  encoder block ....... src/models/KerasLayers.py:696-723 (conv, Dropout, conv, MaxPool)
  decoder block ....... src/models/KerasLayers.py:726-777 (UpSampling2D nearest -> Conv3x3
                        ReLU (no BN) -> Concatenate([deconv, skip]) -> conv, Dropout, conv)
  dropout schedule .... src/models/Unets.py:105-106, :813, :832
  optimizer ........... src/models/ModelUtils.py:75-118 (Adam(lr) Keras defaults)
  loss ................ src/models/Loss_and_metrics.py:6 (mse), :40-89 (loss_with_zero_mask)
  data parallel ....... src/models/Unets.py:70-75 (MirroredStrategy, per-replica BN)

What IS pinned: the structural known answers of the reference's own model.summary()
(notebooks/Train/Train_tests.ipynb cell 9): 8,641,730 params / 8,635,842 trainable /
5,888 non-trainable and the per-layer counts (tests/golden/unet_summary.json), plus a
float64 finite-difference check of the backward pass and an independent numpy
restatement of every op used here (tests/test_oracle_unet.py).

Arithmetic runs in torch CPU ops (float32 or float64); autograd provides the backward
pass, and `manual_block_backward` restates the fused-block formulas the kernels use.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3        # Keras BatchNormalization default epsilon  [TF-2.3]
BN_MOMENTUM = 0.99   # Keras BatchNormalization default momentum [TF-2.3]


# --------------------------------------------------------------------------------------
# configuration (mirrors the config.get(...) defaults of Unets.py:77-106)
# --------------------------------------------------------------------------------------
@dataclass
class NetCfg:
    H: int
    W: int
    in_ch: int
    classes: int
    depth: int
    filters: int
    batch_norm: bool
    bn_first: bool
    use_upsample: bool
    dropouts: List[float]      # encoder level l uses dropouts[l]; decoder pops from the back
    drop_mid: float


def cfg_from_config(config: dict) -> NetCfg:
    dim = config.get('DIM', [224, 224])
    depth = config.get('DEPTH', 4)
    d1 = config.get('DROPOUT_MIN', 0.3)
    d3 = config.get('DROPOUT_MAX', 0.5)
    # Unets.py:105-106 rounds np.float64 values -> numpy's round (x*10, rint, /10), not Python's
    drops = [float(np.round(np.float64(i), 1)) for i in np.linspace(d1, d3, depth)]
    return NetCfg(H=dim[0], W=dim[1], in_ch=config.get('IMG_CHANNELS', 1),
                  classes=config.get('MASK_CLASSES', 3), depth=depth,
                  filters=config.get('FILTERS', 16),
                  batch_norm=bool(config.get('BATCH_NORMALISATION', False)),
                  bn_first=bool(config.get('BN_FIRST', False)),
                  # Unets.py:86 default is the *string* 'False', which is truthy at KerasLayers.py:753
                  use_upsample=bool(config.get('USE_UPSAMPLE', 'False')),
                  dropouts=drops, drop_mid=float(d3))


@dataclass
class ConvSpec:
    name: str
    cin: int
    cout: int
    k: int          # 3 or 1
    bn: bool        # followed by BatchNormalization
    act: str        # 'relu' | 'sigmoid'
    transposed: bool = False   # Conv2DTranspose(k, strides=2, 'same') instead of UpSampling2D + Conv2D (KerasLayers.py:762-765)


def layer_specs(cfg: NetCfg) -> List[ConvSpec]:
    """Conv layers in Keras creation order (Unets.py:786-836, :128)."""
    specs: List[ConvSpec] = []
    c_in, f = cfg.in_ch, cfg.filters
    for l in range(cfg.depth):
        specs.append(ConvSpec(f'enc{l}.conv_a', c_in, f, 3, cfg.batch_norm, 'relu'))
        specs.append(ConvSpec(f'enc{l}.conv_b', f, f, 3, cfg.batch_norm, 'relu'))
        c_in, f = f, f * 2
    specs.append(ConvSpec('mid.conv_a', c_in, f, 3, cfg.batch_norm, 'relu'))
    specs.append(ConvSpec('mid.conv_b', f, f, 3, cfg.batch_norm, 'relu'))
    c_low = f
    for l in range(cfg.depth):
        f //= 2
        specs.append(ConvSpec(f'dec{l}.upconv', c_low, f, 3, False, 'relu', transposed=not cfg.use_upsample))
        specs.append(ConvSpec(f'dec{l}.conv_a', 2 * f, f, 3, cfg.batch_norm, 'relu'))
        specs.append(ConvSpec(f'dec{l}.conv_b', f, f, 3, cfg.batch_norm, 'relu'))
        c_low = f
    specs.append(ConvSpec('head', c_low, cfg.classes, 1, False, 'sigmoid'))
    return specs


def weight_shapes(cfg: NetCfg) -> List[Tuple[str, Tuple[int, ...]]]:
    """model.get_weights() order [TF-2.3]: Conv2D -> kernel HWIO, bias; BN -> gamma, beta,
    moving_mean, moving_variance."""
    out = []
    for s in layer_specs(cfg):
        # Conv2D kernels are HWIO; Conv2DTranspose kernels are (kh, kw, OUT, IN) [TF-2.3]
        out.append((s.name + '/kernel', (s.k, s.k, s.cout, s.cin) if s.transposed else (s.k, s.k, s.cin, s.cout)))
        out.append((s.name + '/bias', (s.cout,)))
        if s.bn:
            for n in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
                out.append((s.name + '/bn/' + n, (s.cout,)))
    return out


def count_params(cfg: NetCfg) -> Tuple[int, int, int]:
    total = train = 0
    for n, shp in weight_shapes(cfg):
        k = int(np.prod(shp))
        total += k
        if not n.endswith(('moving_mean', 'moving_variance')):
            train += k
    return total, train, total - train


def init_weights(cfg: NetCfg, seed: int = 1234, randomize_bn: bool = False) -> List[np.ndarray]:
    """he_normal 3x3 kernels (truncated normal +-2 sigma, stddev sqrt(2/fan_in)/0.87962566),
    glorot_uniform head, zero biases, BN gamma=1 beta=0 mean=0 var=1 [TF-2.3].
    randomize_bn=True perturbs BN tensors so inference tests exercise them."""
    rng = np.random.default_rng(seed)
    ws: List[np.ndarray] = []
    for s in layer_specs(cfg):
        # Keras computes fans from the kernel SHAPE (fan_in = shape[-2] * receptive field): for the (kh, kw, out, in)
        # kernel of Conv2DTranspose that is the OUTPUT channel count [TF-2.3 init_ops._compute_fans]
        shape = (s.k, s.k, s.cout, s.cin) if s.transposed else (s.k, s.k, s.cin, s.cout)
        fan_in = s.k * s.k * shape[2]
        fan_out = s.k * s.k * shape[3]
        if s.name == 'head':
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            k = rng.uniform(-lim, lim, size=shape)
        else:
            std = math.sqrt(2.0 / fan_in) / 0.87962566103423978
            k = rng.standard_normal(size=shape)
            bad = np.abs(k) > 2.0
            while bad.any():
                k[bad] = rng.standard_normal(size=int(bad.sum()))
                bad = np.abs(k) > 2.0
            k = k * std
        ws.append(k.astype(np.float32))
        b = np.zeros(s.cout, np.float32)
        if randomize_bn:
            b = (0.05 * rng.standard_normal(s.cout)).astype(np.float32)
        ws.append(b)
        if s.bn:
            if randomize_bn:
                ws.append(rng.uniform(0.5, 1.5, s.cout).astype(np.float32))
                ws.append((0.1 * rng.standard_normal(s.cout)).astype(np.float32))
                ws.append((0.1 * rng.standard_normal(s.cout)).astype(np.float32))
                ws.append(rng.uniform(0.5, 1.5, s.cout).astype(np.float32))
            else:
                ws.append(np.ones(s.cout, np.float32))
                ws.append(np.zeros(s.cout, np.float32))
                ws.append(np.zeros(s.cout, np.float32))
                ws.append(np.ones(s.cout, np.float32))
    return ws


# --------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------
class _Params:
    """Walks a flat Keras-ordered weight list."""

    def __init__(self, tensors: Sequence[torch.Tensor]):
        self.t = list(tensors)
        self.i = 0

    def take(self, n):
        r = self.t[self.i:self.i + n]
        self.i += n
        return r


def _conv(x, kernel_hwio, bias):
    # Conv2D 'same', stride 1: cross-correlation, zero pad (Appendix C.1). x is NCHW.
    w = kernel_hwio.permute(3, 2, 0, 1)
    pad = kernel_hwio.shape[0] // 2
    return F.conv2d(x, w, bias, stride=1, padding=pad)


def _tconv(x, kernel_hwoi, bias):
    """Conv2DTranspose(kernel 3, strides 2, padding 'same') [TF-2.3]: the gradient w.r.t. the input of the forward
    convolution (2h -> h, stride 2, SAME: total padding max((h-1)*2 + 3 - 2h, 0) = 1, all of it at the END since
    pad_before = total // 2 = 0), i.e. out[o] = sum_{2i + k = o} x[i] * w[k] cropped to the first 2h rows / columns.
    torch's conv_transpose2d is the same scatter (weight [in, out, kh, kw]) on a (2h + 1)-long output."""
    h, w = x.shape[2], x.shape[3]
    y = F.conv_transpose2d(x, kernel_hwoi.permute(3, 2, 0, 1), bias, stride=2)
    return y[:, :, :2 * h, :2 * w]


def _bn(x, gamma, beta, mm, mv, training, new_stats, name):
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        n = x.shape[0] * x.shape[2] * x.shape[3]
        with torch.no_grad():
            unb = var * (n / max(n - 1, 1))
            new_stats[name] = ((BN_MOMENTUM * mm + (1 - BN_MOMENTUM) * mean).detach(),
                               (BN_MOMENTUM * mv + (1 - BN_MOMENTUM) * unb).detach())
    else:
        mean, var = mm, mv
    inv = torch.rsqrt(var + BN_EPS)
    return (x - mean[None, :, None, None]) * (inv * gamma)[None, :, None, None] + beta[None, :, None, None]


class _RoundFwd(torch.autograd.Function):
    """bf16-storage emulation: round a stored activation, straight-through gradient."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    """bf16-storage emulation: identity forward, round the gradient flowing back (a stored dz / dx)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


_STORAGE = {'mode': 'fp32'}


def _block(x, p: _Params, spec: ConvSpec, cfg: NetCfg, training, new_stats, acts):
    if _STORAGE['mode'] == 'bf16':
        return _block_bf16(x, p, spec, cfg, training, new_stats, acts)
    k, b = p.take(2)
    z = _tconv(x, k, b) if spec.transposed else _conv(x, k, b)
    if spec.bn:
        g, be, mm, mv = p.take(4)
        if cfg.bn_first:
            y = torch.relu(_bn(z, g, be, mm, mv, training, new_stats, spec.name))
            acts[spec.name + '/a'] = z
        else:
            a = torch.relu(z)
            acts[spec.name + '/a'] = a
            y = _bn(a, g, be, mm, mv, training, new_stats, spec.name)
    else:
        y = torch.relu(z)
        acts[spec.name + '/a'] = y
    acts[spec.name + '/y'] = y
    return y


def _block_bf16(x, p: _Params, spec: ConvSpec, cfg: NetCfg, training, new_stats, acts):
    """Same block with the device path's bf16 storage points restated: tensor-core operands (weights of
    every conv but the Cin=in_ch first layer), the stored relu(conv) `a`, the stored block output `y`,
    and in backward the stored dz (gradient at the conv output) and dx (gradient at the conv input).
    Accumulation stays fp32, as in TMEM. Used only to CALIBRATE what bf16 storage itself does to the
    gradients of this ill-conditioned (BatchNorm, random init) problem; the fp32 oracle stays the spec."""
    assert not cfg.bn_first
    k, b = p.take(2)
    first = spec.cin == cfg.in_ch and spec.name == 'enc0.conv_a'
    if not first:
        k = _RoundFwd.apply(k)
        x = _RoundBwd.apply(x)
    z = _RoundBwd.apply(_tconv(x, k, b) if spec.transposed else _conv(x, k, b))
    a = _RoundFwd.apply(torch.relu(z))
    acts[spec.name + '/a'] = a
    if spec.bn:
        g, be, mm, mv = p.take(4)
        y = _RoundFwd.apply(_bn(a, g, be, mm, mv, training, new_stats, spec.name))
    else:
        y = a
    acts[spec.name + '/y'] = y
    return y


def _phase_taps(parity, r):
    """3x3 tap indices that fall on low-resolution neighbour r (0 / 1) of an output row / column of this parity."""
    return ((0,), (1, 2))[r] if parity == 0 else ((0, 1), (2,))[r]


def phased_upconv_eligible(h, w, cin, cout):
    """Shapes the device path computes with the phase-decomposed up-convolution (conv_halo.cuh); h, w = low resolution."""
    return h >= 1 and w >= 1 and cin % 64 == 0 and (cout % 64 == 0 or cout == 32)


def _upconv_phased_bf16(x_low, p: _Params, spec: ConvSpec, acts):
    """Calibration restatement of the device's phase-decomposed up-convolution: UpSampling2D(2) -> Conv3x3 is the same
    linear map as four 2x2-tap convolutions on the low-resolution tensor (one per output parity (a, b)) whose weights are
    sums of the 3x3 taps that coincide after nearest-neighbour up-sampling (KerasLayers.py:756-759). The device sums in
    fp32 and rounds the SUM to bf16 once; everything else is as in _block_bf16."""
    k, b = p.take(2)
    x = _RoundBwd.apply(x_low)
    h, w = x.shape[2], x.shape[3]
    xp = F.pad(x, (1, 1, 1, 1))
    rows = []
    for a in range(2):
        cols = []
        for bb in range(2):
            kp = torch.stack([torch.stack([sum(k[ky, kx] for ky in _phase_taps(a, r) for kx in _phase_taps(bb, s_))
                                           for s_ in range(2)]) for r in range(2)])          # [2, 2, Cin, C]
            kp = _RoundFwd.apply(kp)
            cols.append(F.conv2d(xp, kp.permute(3, 2, 0, 1), b)[:, :, a:a + h, bb:bb + w])
        rows.append(torch.stack(cols, dim=-1))                  # [B, C, h, w, 2]
    z = torch.stack(rows, dim=3).reshape(x.shape[0], k.shape[3], 2 * h, 2 * w)   # [B, C, h, 2, w, 2] -> [B, C, 2h, 2w]
    a_ = _RoundFwd.apply(torch.relu(_RoundBwd.apply(z)))
    acts[spec.name + '/a'] = a_
    acts[spec.name + '/y'] = a_
    return a_


def _dropout(x, rate, training, masks, name):
    if not training or rate == 0.0:
        return x
    if masks is None or name not in masks:
        raise ValueError('oracle dropout needs an explicit keep-mask for %s (TF RNG cannot be matched)' % name)
    m = masks[name]                                  # NHWC keep mask (0/1)
    m = torch.as_tensor(m, dtype=x.dtype).permute(0, 3, 1, 2)
    return x * m / (1.0 - rate)


def forward_torch(cfg: NetCfg, params: Sequence[torch.Tensor], x_nchw: torch.Tensor, training: bool,
                  dropout_masks: Optional[Dict[str, np.ndarray]] = None):
    """Returns (heatmap NCHW, acts dict, new BN moving stats dict)."""
    specs = {s.name: s for s in layer_specs(cfg)}
    p = _Params(params)
    acts: Dict[str, torch.Tensor] = {}
    new_stats: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
    skips = []
    h = x_nchw
    for l in range(cfg.depth):
        h = _block(h, p, specs[f'enc{l}.conv_a'], cfg, training, new_stats, acts)
        h = _dropout(h, cfg.dropouts[l], training, dropout_masks, f'enc{l}')
        h = _block(h, p, specs[f'enc{l}.conv_b'], cfg, training, new_stats, acts)
        skips.append(h)
        h = F.max_pool2d(h, 2, 2)                     # MaxPooling2D((2,2)) valid (Appendix C.5)
    h = _block(h, p, specs['mid.conv_a'], cfg, training, new_stats, acts)
    h = _dropout(h, cfg.drop_mid, training, dropout_masks, 'mid')
    h = _block(h, p, specs['mid.conv_b'], cfg, training, new_stats, acts)
    drops = list(cfg.dropouts)
    for l in range(cfg.depth):
        skip = skips.pop()
        su = specs[f'dec{l}.upconv']
        if not cfg.use_upsample:
            u = _block(h, p, su, cfg, training, new_stats, acts)      # Conv2DTranspose straight on the low-resolution tensor
        elif (_STORAGE['mode'] == 'bf16' and _STORAGE.get('phased_up')
                and phased_upconv_eligible(h.shape[2], h.shape[3], su.cin, su.cout)):
            u = _upconv_phased_bf16(h, p, su, acts)
        else:
            up = F.interpolate(h, scale_factor=2, mode='nearest')       # UpSampling2D (Appendix C.6)
            u = _block(up, p, su, cfg, training, new_stats, acts)
        h = torch.cat([u, skip], dim=1)               # Concatenate([deconv, skip]) KerasLayers.py:767
        h = _block(h, p, specs[f'dec{l}.conv_a'], cfg, training, new_stats, acts)
        h = _dropout(h, drops.pop(), training, dropout_masks, f'dec{l}')
        h = _block(h, p, specs[f'dec{l}.conv_b'], cfg, training, new_stats, acts)
    k, b = p.take(2)
    if _STORAGE['mode'] == 'bf16':
        h = _RoundBwd.apply(h)
    logits = _conv(h, k, b)
    acts['head/logits'] = logits
    heat = torch.sigmoid(logits)                      # Unets.py:128
    assert p.i == len(p.t), 'weight list length mismatch'
    return heat, acts, new_stats


def _to_params(weights: Sequence[np.ndarray], dtype, requires_grad=False):
    ps = []
    names = [n for n, _ in []]
    for w in weights:
        t = torch.tensor(np.asarray(w), dtype=dtype)
        ps.append(t)
    return ps


def trainable_mask(cfg: NetCfg) -> List[bool]:
    return [not n.endswith(('moving_mean', 'moving_variance')) for n, _ in weight_shapes(cfg)]


def predict(cfg: NetCfg, weights: Sequence[np.ndarray], x_nhwc: np.ndarray, dtype=torch.float32,
            return_acts: bool = False):
    """model.predict semantics: BN moving stats, no dropout (Appendix C.13)."""
    with torch.no_grad():
        ps = _to_params(weights, dtype)
        x = torch.tensor(x_nhwc, dtype=dtype).permute(0, 3, 1, 2)
        heat, acts, _ = forward_torch(cfg, ps, x, training=False)
    out = heat.permute(0, 2, 3, 1).contiguous().numpy()
    if return_acts:
        return out, {k: v.permute(0, 2, 3, 1).contiguous().numpy() for k, v in acts.items()}
    return out


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
def inplane_weights(H: int, W: int) -> np.ndarray:
    """Loss_and_metrics.py:62-69: temp[i:-i, i:-i] = linspace(0,100,xy//2)[i] -> concentric square
    ramp (i = 0 selects an empty slice, so the border keeps 0). Generalised to H x W with the
    ramp length taken from the first dimension exactly like the reference (x_shape // 2)."""
    temp = np.zeros((H, W))
    dist = np.linspace(0, 100, H // 2)
    for i, l in enumerate(dist):
        temp[i:-i, i:-i] = l
    return temp.astype(np.float32)


def loss_torch(heat_nchw, target_nchw, kind='mse', mask_smaller_than=0.01, weights_hw=None, eps=1e-7, w_bce=1.0,
               w_dice=1.0):
    """'mse': mean((t-p)^2, axis=-1) then Keras SUM_OVER_BATCH_SIZE = global mean (Appendix C.11/14).
    'masked' / 'weighted': loss_with_zero_mask (Loss_and_metrics.py:40-89). The reference squeezes
    the mask on axis -1 (C must be 1); for C>1 the per-pixel mask is any_c(t > thr) -- an
    EXTENSION that reduces to the reference for C=1."""
    if kind == 'bce_dice':
        # BceDiceLoss.__call__ (Loss_and_metrics.py:208-228; bce_dice_loss :231-245 is the same with w_bce = 0.5):
        #   w_bce * keras.binary_crossentropy(t, p) - w_dice * dice_coef(t, p)          (per pixel, dice a batch scalar)
        # keras binary_crossentropy on probabilities [TF-2.3 backend]: p clipped to [eps, 1-eps], then
        #   -mean_c( t * log(p + eps) + (1 - t) * log(1 - p + eps) );  dice_coef (:165-171): smooth = 1, raw p.
        # Reduction: mean over (B, H, W) -- the value Keras logs.  [TF-2.3 semantics, unpinned]: the class overrides
        # __call__, so Keras differentiates the SUM of the unreduced tensor; that is this gradient times B*H*W, a
        # constant factor Adam's normalisation removes (up to its epsilon).
        p = heat_nchw.clamp(eps, 1 - eps)
        bce = -(target_nchw * torch.log(p + eps) + (1 - target_nchw) * torch.log(1 - p + eps)).mean(dim=1)
        inter = (target_nchw * heat_nchw).sum()
        dice = (2.0 * inter + 1.0) / (target_nchw.sum() + heat_nchw.sum() + 1.0)
        return (w_bce * bce - w_dice * dice).mean()
    per_px = ((target_nchw - heat_nchw) ** 2).mean(dim=1)          # [B,H,W]
    if kind == 'mse':
        return per_px.mean()
    mask = (target_nchw > mask_smaller_than).any(dim=1).to(per_px.dtype)
    if kind == 'masked':
        return (per_px * mask).mean()
    if kind == 'weighted':
        w = torch.as_tensor(weights_hw, dtype=per_px.dtype)
        return ((per_px * mask) * w[None] + eps).mean()
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# training step (fwd, loss, bwd) and Adam
# --------------------------------------------------------------------------------------
def train_grads(cfg: NetCfg, weights: Sequence[np.ndarray], x_nhwc, t_nhwc, dtype=torch.float32,
                loss_kind='mse', dropout_masks=None, return_acts=False, weights_hw=None, storage='fp32',
                loss_params=None, phased_up=False):
    """One replica's forward + loss + backward. Returns dict(loss, heat, grads (Keras order,
    None for non-trainable), new_stats, [acts, act_grads]). storage='bf16' restates the device path's
    bf16 storage points (see _block_bf16); phased_up additionally restates its pre-summed up-convolution weights."""
    _STORAGE['mode'] = storage
    _STORAGE['phased_up'] = bool(phased_up)
    try:
        return _train_grads(cfg, weights, x_nhwc, t_nhwc, dtype, loss_kind, dropout_masks, return_acts, weights_hw,
                            loss_params or {})
    finally:
        _STORAGE['mode'] = 'fp32'
        _STORAGE['phased_up'] = False


def _train_grads(cfg, weights, x_nhwc, t_nhwc, dtype, loss_kind, dropout_masks, return_acts, weights_hw, loss_params):
    ps = _to_params(weights, dtype)
    tm = trainable_mask(cfg)
    for p_, t_ in zip(ps, tm):
        p_.requires_grad_(t_)
    x = torch.tensor(x_nhwc, dtype=dtype).permute(0, 3, 1, 2)
    t = torch.tensor(t_nhwc, dtype=dtype).permute(0, 3, 1, 2)
    heat, acts, new_stats = forward_torch(cfg, ps, x, training=True, dropout_masks=dropout_masks)
    if return_acts:
        for v in acts.values():
            if v.requires_grad:
                v.retain_grad()
    loss = loss_torch(heat, t, loss_kind, weights_hw=weights_hw, **loss_params)
    loss.backward()
    grads = [p_.grad.numpy().copy() if t_ else None for p_, t_ in zip(ps, tm)]
    out = dict(loss=float(loss.detach()), heat=heat.detach().permute(0, 2, 3, 1).contiguous().numpy(),
               grads=grads,
               new_stats={k: (a.numpy().copy(), b.numpy().copy()) for k, (a, b) in new_stats.items()})
    if return_acts:
        out['acts'] = {k: v.detach().permute(0, 2, 3, 1).contiguous().numpy() for k, v in acts.items()}
        out['act_grads'] = {k: v.grad.permute(0, 2, 3, 1).contiguous().numpy()
                            for k, v in acts.items() if v.grad is not None}
    return out


def apply_new_stats(cfg: NetCfg, weights: List[np.ndarray], new_stats) -> List[np.ndarray]:
    out = list(weights)
    names = [n for n, _ in weight_shapes(cfg)]
    for i, n in enumerate(names):
        if n.endswith('/bn/moving_mean'):
            out[i] = new_stats[n[:-len('/bn/moving_mean')]][0].astype(np.float32)
        elif n.endswith('/bn/moving_variance'):
            out[i] = new_stats[n[:-len('/bn/moving_variance')]][1].astype(np.float32)
    return out


class Adam:
    """tf.keras.optimizers.Adam(lr) defaults: b1 .9, b2 .999, eps 1e-7, epsilon-hat form
    p -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)   (Appendix C.12; ModelUtils.py:107)."""

    def __init__(self, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.t = 0
        self.m = None
        self.v = None

    def step(self, weights: List[np.ndarray], grads: List[Optional[np.ndarray]]) -> List[np.ndarray]:
        if self.m is None:
            self.m = [None if g is None else np.zeros_like(w, np.float64) for w, g in zip(weights, grads)]
            self.v = [None if g is None else np.zeros_like(w, np.float64) for w, g in zip(weights, grads)]
        self.t += 1
        a = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        out = []
        for i, (w, g) in enumerate(zip(weights, grads)):
            if g is None:
                out.append(w)
                continue
            g = g.astype(np.float64)
            self.m[i] = self.b1 * self.m[i] + (1 - self.b1) * g
            self.v[i] = self.b2 * self.v[i] + (1 - self.b2) * g * g
            out.append((w.astype(np.float64) - a * self.m[i] / (np.sqrt(self.v[i]) + self.eps)).astype(np.float32))
        return out


class SGD:
    """tf.keras.optimizers.SGD [TF-2.3] (ModelUtils.py:109-111: SGD(lr, nesterov=True); KerasCallbacks.py:295: SGD()):
    v = momentum * v - lr * g;  w += momentum * v - lr * g if nesterov else v.  float64 state, float32 weights."""

    def __init__(self, lr=0.01, momentum=0.0, nesterov=False):
        self.lr, self.momentum, self.nesterov = lr, momentum, nesterov
        self.v: Optional[List[Optional[np.ndarray]]] = None

    def step(self, weights: List[np.ndarray], grads: List[Optional[np.ndarray]]) -> List[np.ndarray]:
        if self.v is None:
            self.v = [None if g is None else np.zeros(g.shape, np.float64) for g in grads]
        out = []
        for i, (w, g) in enumerate(zip(weights, grads)):
            if g is None:
                out.append(w)
                continue
            g = g.astype(np.float64)
            self.v[i] = self.momentum * self.v[i] - self.lr * g
            upd = self.momentum * self.v[i] - self.lr * g if self.nesterov else self.v[i]
            out.append((w.astype(np.float64) + upd).astype(np.float32))
        return out


def data_parallel_grads(cfg, weights, x_nhwc, t_nhwc, world: int, dtype=torch.float32, loss_kind='mse'):
    """MirroredStrategy semantics (Unets.py:70-75; Appendix C.11): each replica runs its shard with
    per-replica BN statistics; per-replica loss is divided by the GLOBAL batch and gradients are
    summed == mean of per-replica (local-mean-loss) gradients. Moving stats averaged on read."""
    B = x_nhwc.shape[0]
    assert B % world == 0
    per = B // world
    outs = [train_grads(cfg, weights, x_nhwc[r * per:(r + 1) * per], t_nhwc[r * per:(r + 1) * per],
                        dtype, loss_kind) for r in range(world)]
    grads = []
    for i in range(len(weights)):
        if outs[0]['grads'][i] is None:
            grads.append(None)
        else:
            grads.append(sum(o['grads'][i] for o in outs) / world)
    loss = sum(o['loss'] for o in outs) / world
    stats = {k: (sum(o['new_stats'][k][0] for o in outs) / world, sum(o['new_stats'][k][1] for o in outs) / world)
             for k in outs[0]['new_stats']}
    return dict(loss=loss, grads=grads, new_stats=stats, per_rank=outs)


# --------------------------------------------------------------------------------------
# manual restatement of the fused-block backward (Appendix C.15) -- what the kernels implement
# --------------------------------------------------------------------------------------
def manual_block_backward(a, dy, gamma, eps=BN_EPS):
    """a = relu(z) [N,C] (pixels x channels), dy = dL/dy with y = gamma*ahat+beta.
    Returns dz, dgamma, dbeta."""
    mu = a.mean(0)
    var = a.var(0)
    r = 1.0 / np.sqrt(var + eps)
    ahat = (a - mu) * r
    dbeta = dy.sum(0)
    dgamma = (dy * ahat).sum(0)
    n = a.shape[0]
    da = gamma * r * (dy - dbeta / n - ahat * dgamma / n)
    dz = da * (a > 0)
    return dz, dgamma, dbeta
