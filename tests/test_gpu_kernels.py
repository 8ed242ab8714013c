"""GPU parity tests of the individual kernels, called through the C ABI (librvip_b200.so)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CONV_SHAPES = [  # B, H, W, C0, C1, Cout
    (2, 16, 16, 64, 0, 64),
    (2, 32, 32, 32, 0, 32),
    (1, 16, 16, 32, 32, 64),      # concat, KC=32
    (2, 8, 8, 128, 0, 256),
    (3, 4, 4, 64, 0, 512),        # several images per tile, batch not a multiple of NB, two N tiles
    (1, 24, 40, 64, 0, 32),       # non power-of-two image: partial tiles
    (2, 16, 16, 64, 64, 128),     # concat, KC=64
    (1, 128, 128, 32, 0, 32),     # row tiles of 128 pixels
]


ROW_SHAPES = [  # W % 128 == 0: eligible for the row-tiled kernels
    (1, 8, 128, 32, 0, 32),
    (2, 8, 256, 64, 0, 32),
    (1, 8, 128, 32, 32, 32),
    (1, 8, 128, 64, 0, 64),
    (1, 4, 128, 128, 0, 64),      # weights streamed per stage (not resident)
    (1, 6, 128, 32, 0, 32),       # H % 4 != 0 -> R = 2
    (2, 16, 128, 64, 64, 64),
]
ALL_CASES = [('tc', s) for s in CONV_SHAPES] + [('row', s) for s in ROW_SHAPES]
HALO_SHAPES = [  # W % 16 == 0: eligible for the halo-staged wgrad kernel
    (2, 16, 16, 64, 0, 128),      # CIC=32, BN=128, TW=16
    (2, 32, 32, 128, 0, 256),     # two N tiles, TW=32
    (1, 64, 64, 64, 0, 64),       # CIC=64 (taps paired), BN=64
    (2, 32, 32, 64, 64, 64),      # concat, CIC=64
    (1, 32, 32, 32, 32, 64),      # concat on a 32-channel boundary -> CIC=32, BN=64
    (1, 24, 48, 64, 0, 32),       # H not a multiple of the tile height, BN=32
    (3, 16, 16, 32, 0, 32),       # CIC=32, BN=32
    (1, 8, 16, 256, 0, 128),      # image lower than the nominal tile
    (2, 128, 128, 128, 0, 64),    # level-1 shape of the bench network
    (2, 28, 28, 64, 0, 64),       # image sizes off the 16-pixel grid (the reference trains at 224: levels 56 / 28 / 14):
    (1, 14, 14, 128, 0, 128),     #   columns / rows past the image are zero-filled by TMA
    (1, 56, 40, 32, 32, 64),
]
WGRAD_CASES = ALL_CASES + [('halo', s) for s in HALO_SHAPES]
CONV_HALO_SHAPES = [  # H, W % 16 == 0, channels % 64 == 0: eligible for the halo-staged conv kernel
    (2, 16, 16, 64, 0, 64),       # one block per image, BN=64
    (2, 32, 32, 128, 0, 256),     # BN=256 (single TMEM buffer), 4 blocks per image
    (1, 32, 48, 64, 64, 128),     # concat, BN=128, non-square
    (3, 16, 16, 256, 0, 512),     # two N tiles, persistent loop over several tiles per CTA
    (40, 16, 16, 64, 0, 128),     # more tiles than SMs: double-buffered accumulators wrap around
    (2, 16, 16, 128, 0, 192),     # Cout = 3 x 64
    (2, 24, 40, 64, 0, 64),       # image sizes off the 16-pixel grid: partial blocks, statistics skip the outside rows
    (1, 14, 14, 128, 0, 128),     # image smaller than one block (mid level of a 224 x 224 input)
    (3, 28, 28, 64, 64, 64),      # concat + partial blocks
]
ROW_EDGE_SHAPES = [  # rows that are not a whole number of 128-pixel tiles (the reference's 224 x 224 input and its 112 level)
    (1, 8, 224, 32, 0, 32),
    (2, 4, 96, 64, 0, 64),
    (1, 8, 224, 32, 32, 32),
    (1, 4, 112, 64, 0, 64),
]
CONV_CASES = ALL_CASES + [('halo', s) for s in CONV_HALO_SHAPES] + [('row', s) for s in ROW_EDGE_SHAPES]


def _rand_bf16(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device='cuda') * scale).to(torch.bfloat16)


@pytest.mark.parametrize('variant,shape', CONV_CASES)
def test_conv_tc_forward_bias_relu_stats(variant, shape):
    from tests import gpu_util as U
    conv = {'tc': U.conv_tc, 'row': U.conv_row, 'halo': U.conv_halo}[variant]
    B, H, W, C0, C1, N = shape
    g = torch.Generator(device='cuda').manual_seed(1 + sum(shape))
    x0 = _rand_bf16((B, H, W, C0), g)
    x1 = _rand_bf16((B, H, W, C1), g) if C1 else None
    w = torch.randn((3, 3, C0 + C1, N), generator=g, device='cuda') * (2.0 / (9 * (C0 + C1))) ** 0.5
    bias = torch.randn(N, generator=g, device='cuda') * 0.1
    out, _, stats = conv(x0, x1, U.pack_fwd(w), bias, N, mode=0, want_stats=True)
    xin = x0 if x1 is None else torch.cat([x0, x1], dim=3)
    ref = U.ref_conv(xin, w, bias, relu=True)
    assert torch.isfinite(out.float()).all()
    # bf16 output rounding: <= 2^-8 relative per element
    assert U.max_err(out, ref) <= 1e-2 * float(ref.abs().max()) + 1e-3
    assert U.rel_err(out, ref) < 4e-3
    rb = out.double()
    assert torch.allclose(stats[:N], rb.sum(dim=(0, 1, 2)), rtol=2e-3, atol=1e-2)
    assert torch.allclose(stats[N:], (rb * rb).sum(dim=(0, 1, 2)), rtol=2e-3, atol=1e-2)


@pytest.mark.parametrize('variant,shape', CONV_CASES)
def test_conv_tc_dgrad(variant, shape):
    from tests import gpu_util as U
    conv = {'tc': U.conv_tc, 'row': U.conv_row, 'halo': U.conv_halo}[variant]
    B, H, W, C0, C1, N = shape
    g = torch.Generator(device='cuda').manual_seed(2 + sum(shape))
    dz = _rand_bf16((B, H, W, N), g)
    w = torch.randn((3, 3, C0 + C1, N), generator=g, device='cuda') * (2.0 / (9 * N)) ** 0.5
    dx0, dx1, _ = conv(dz, None, U.pack_dgrad(w), None, C0 + C1, mode=2, out_split=C0 if C1 else C0 + C1)
    x = torch.zeros((B, C0 + C1, H, W), device='cuda', requires_grad=True)
    wt = w.to(torch.bfloat16).float().permute(3, 2, 0, 1)
    y = torch.nn.functional.conv2d(x, wt, padding=1)
    y.backward(dz.float().permute(0, 3, 1, 2))
    ref = x.grad.permute(0, 2, 3, 1)
    got = dx0 if dx1 is None else torch.cat([dx0, dx1], dim=3)
    assert U.rel_err(got, ref) < 4e-3


UPCONV_SHAPES = [  # B, h, w (LOW resolution), Cin, C
    (2, 16, 16, 64, 32),          # two row phases, column phase merged into N = 64 (2 x 3 taps)
    (1, 32, 48, 64, 32),          # non-square, several blocks
    (2, 16, 16, 128, 64),         # four phases of 2 x 2 taps, BN = 64
    (1, 32, 32, 256, 128),        # BN = 128, four K chunks
    (3, 16, 16, 512, 256),        # deepest decoder level of the bench network
    (40, 16, 16, 128, 64),        # more tiles than SMs
    (2, 14, 14, 128, 64),         # low-resolution sizes off the 16-pixel grid (224 x 224 input: 14 / 28 / 56 / 112)
    (1, 28, 28, 64, 32),
    (1, 56, 40, 64, 32),
]


def _phase_sets(p, r):
    """3x3 tap indices (0..2) that land on low-resolution neighbour r of output parity p."""
    return ([0], [1, 2])[r] if p == 0 else ([0, 1], [2])[r]


def _ref_upconv_phased(x_low, w_hwio, bias):
    """Same arithmetic as the device path: taps pre-summed in fp32, ONE rounding to bf16, fp32 accumulation."""
    B, h, w, cin = x_low.shape
    c = w_hwio.shape[3]
    xp = torch.nn.functional.pad(x_low.float().permute(0, 3, 1, 2), (1, 1, 1, 1))
    out = torch.empty((B, c, 2 * h, 2 * w), device=x_low.device)
    for a in range(2):
        for b in range(2):
            k = torch.zeros((2, 2, cin, c), device=x_low.device)
            for r in range(2):
                for s_ in range(2):
                    for ky in _phase_sets(a, r):
                        for kx in _phase_sets(b, s_):
                            k[r, s_] += w_hwio[ky, kx].float()
            k = k.to(torch.bfloat16).float().permute(3, 2, 0, 1)
            o = torch.nn.functional.conv2d(xp, k, bias)
            out[:, :, a::2, b::2] = o[:, :, a:a + h, b:b + w]
    return torch.relu(out).permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize('shape', UPCONV_SHAPES)
def test_upconv_phased_forward(shape):
    """UpSampling2D(2) -> Conv3x3 -> ReLU from the low-resolution tensor (KerasLayers.py:756-759)."""
    from tests import gpu_util as U
    B, h, w, cin, c = shape
    g = torch.Generator(device='cuda').manual_seed(11 + sum(shape))
    x = _rand_bf16((B, h, w, cin), g)
    wt = torch.randn((3, 3, cin, c), generator=g, device='cuda') * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(c, generator=g, device='cuda') * 0.1
    out = U.upconv_halo(0, x, wt, bias)
    assert torch.isfinite(out.float()).all()
    tight = _ref_upconv_phased(x, wt, bias)
    assert U.rel_err(out, tight) < 4e-3                       # bf16 output rounding only
    true = U.ref_upconv(x, wt, bias, relu=True)               # un-summed fp32 weights on the up-sampled tensor
    assert U.rel_err(out, true) < 8e-3                        # + one bf16 rounding of the pre-summed weights


@pytest.mark.parametrize('shape', UPCONV_SHAPES)
def test_upconv_phased_dgrad(shape):
    from tests import gpu_util as U
    B, h, w, cin, c = shape
    g = torch.Generator(device='cuda').manual_seed(12 + sum(shape))
    dz = _rand_bf16((B, 2 * h, 2 * w, c), g)
    wt = torch.randn((3, 3, cin, c), generator=g, device='cuda') * (2.0 / (9 * c)) ** 0.5
    dx = U.upconv_halo(1, dz, wt)
    assert torch.isfinite(dx.float()).all()
    ref = U.ref_upconv(torch.zeros((B, h, w, cin), device='cuda'), wt, None, relu=False, dz=dz)
    assert U.rel_err(dx, ref) < 8e-3


@pytest.mark.parametrize('shape', UPCONV_SHAPES + [(2, 24, 48, 64, 96), (1, 128, 128, 64, 32)])
def test_upconv_phased_wgrad(shape):
    """dW of UpSampling2D(2) -> Conv3x3 from the low-resolution input: eight phase accumulators folded onto the 3x3 taps."""
    from tests import gpu_util as U
    B, h, w, cin, c = shape
    g = torch.Generator(device='cuda').manual_seed(13 + sum(shape))
    x = _rand_bf16((B, h, w, cin), g)
    dz = _rand_bf16((B, 2 * h, 2 * w, c), g)
    dw = U.upconv_wgrad_halo(x, dz)
    wv = torch.zeros((c, cin, 3, 3), device='cuda', requires_grad=True)
    up = torch.nn.functional.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode='nearest')
    y = torch.nn.functional.conv2d(up, wv, padding=1)
    y.backward(dz.float().permute(0, 3, 1, 2))
    ref = wv.grad.permute(2, 3, 1, 0)
    assert U.rel_err(dw, ref) < 2e-3


@pytest.mark.parametrize('shape', UPCONV_SHAPES[:5])
def test_conv2d_transpose_on_the_phase_kernels(shape):
    """USE_UPSAMPLE falsy: Conv2DTranspose(3, strides 2, 'same') -> ReLU (KerasLayers.py:762-765) forward, input gradient
    and kernel gradient on the phase-decomposed kernels (single-tap phase weights), against torch conv_transpose2d."""
    from tests import gpu_util as U
    B, h, w, cin, c = shape
    g = torch.Generator(device='cuda').manual_seed(21 + sum(shape))
    x = _rand_bf16((B, h, w, cin), g)
    wt = torch.randn((3, 3, c, cin), generator=g, device='cuda') * (2.0 / (9 * c)) ** 0.5       # (kh, kw, out, in)
    bias = torch.randn(c, generator=g, device='cuda') * 0.1
    dz = _rand_bf16((B, 2 * h, 2 * w, c), g)
    out = U.upconv_halo(0, x, wt, bias, transposed=True)
    ref, xl, wl = U.ref_tconv(x, wt, bias, relu=True)
    assert torch.isfinite(out.float()).all()
    assert U.rel_err(out, ref.detach().permute(0, 2, 3, 1)) < 4e-3
    lin, xl, wl = U.ref_tconv(x, wt, None, relu=False)
    lin.backward(dz.float().permute(0, 3, 1, 2))
    dx = U.upconv_halo(1, dz, wt, transposed=True)
    assert U.rel_err(dx, xl.grad.permute(0, 2, 3, 1)) < 4e-3
    dw = U.upconv_wgrad_halo(x, dz, transposed=True)
    assert dw.shape == wl.grad.shape and U.rel_err(dw, wl.grad) < 2e-3


@pytest.mark.parametrize('variant,shape', WGRAD_CASES)
def test_wgrad_tc(variant, shape):
    from tests import gpu_util as U
    B, H, W, C0, C1, N = shape
    g = torch.Generator(device='cuda').manual_seed(3 + sum(shape))
    x0 = _rand_bf16((B, H, W, C0), g)
    x1 = _rand_bf16((B, H, W, C1), g) if C1 else None
    dz = _rand_bf16((B, H, W, N), g)
    dw = {'tc': U.wgrad_tc, 'row': U.wgrad_row, 'halo': U.wgrad_halo}[variant](x0, x1, dz)
    xin = (x0 if x1 is None else torch.cat([x0, x1], dim=3)).float().permute(0, 3, 1, 2)
    w = torch.zeros((N, C0 + C1, 3, 3), device='cuda', requires_grad=True)
    y = torch.nn.functional.conv2d(xin, w, padding=1)
    y.backward(dz.float().permute(0, 3, 1, 2))
    ref = w.grad.permute(2, 3, 1, 0)
    assert U.rel_err(dw, ref) < 2e-3


def test_extract_matches_reference_golden(golden_dir):
    """Device kernel vs outputs of the reference's own get_ip_from_rvip_mask_3d (golden) and vs the oracle."""
    from cmr_landmark_detection_b200.extract import extract_device, label_map_device, points_from_stats
    from oracle import extract_ref as ex
    gold = np.load(os.path.join(golden_dir, 'extract_golden.npz'))
    for key in gold['names']:
        heat = gold[key + '/heat']
        r = extract_device(torch.from_numpy(heat).cuda())
        yx, cnt = r['yx'].cpu().numpy(), r['count'].cpu().numpy()
        lab = label_map_device(torch.from_numpy(heat).cuda()).cpu().numpy()
        assert np.array_equal(lab, gold[key + '/labels']), key
        for both in (True, False):
            a, b = points_from_stats(yx, cnt, keepdim=True, both_only=both)
            for got, want in ((a, gold[key + '/ant_both%d' % both]), (b, gold[key + '/inf_both%d' % both])):
                for p, q in zip(got, want):
                    if np.isnan(q[0]):
                        assert p is None
                    else:
                        assert abs(p[0] - q[0]) <= 1e-4 and abs(p[1] - q[1]) <= 1e-4     # tolerance: 1e-4 px
        count, srow, scol, amax, vmax = ex.extract_stats(heat)
        assert np.array_equal(cnt, count)
        finite = ~np.isnan(heat).any(axis=(1, 2))
        assert np.array_equal(r['argmax'].cpu().numpy()[finite], amax[finite])       # bit-exact
        assert np.array_equal(r['maxv'].cpu().numpy()[finite], vmax[finite])


def test_extract_full_size_properties():
    """BASELINE config 4 size (16 x 256 x 256 x 2): argmax bit-exact vs numpy, centroid of a shifted volume
    shifts by exactly the shift, empty volume -> all absent."""
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.extract import extract_device
    vol = synth.make_volume_heat(16, 256, 256, seed=3)
    r = extract_device(torch.from_numpy(vol).cuda())
    am = vol.reshape(16, -1, 2).argmax(axis=1)
    assert np.array_equal(r['argmax'].cpu().numpy(), am)
    rolled = np.roll(vol, shift=(3, -5), axis=(1, 2))
    r2 = extract_device(torch.from_numpy(rolled).cuda())
    d = (r2['yx'] - r['yx']).cpu().numpy()
    ok = np.isfinite(d[..., 0])
    inside = ok & (np.abs(r['yx'].cpu().numpy()[..., 0] - 128) < 90) & (np.abs(r['yx'].cpu().numpy()[..., 1] - 128) < 90)
    # noise pixels never exceed thr, so a blob away from the border translates rigidly
    assert np.allclose(d[inside][:, 0], 3.0, atol=1e-9) and np.allclose(d[inside][:, 1], -5.0, atol=1e-9)
    z = extract_device(torch.zeros((2, 256, 256, 2), device='cuda'))
    assert int(z['count'].sum()) == 0 and bool(torch.isnan(z['yx']).all())
    assert np.array_equal(z['argmax'].cpu().numpy(), np.zeros((2, 2), np.int32))


def test_reference_signature_wrappers():
    from src.models.evaluate_cv import get_ip_from_rvip_mask_3d, get_mean_rvip_2d
    m = np.zeros((32, 32), np.uint8)
    m[5:8, 10:13] = 1
    m[20:22, 4:9] = 2
    assert get_mean_rvip_2d(m) == ([6.0, 11.0], [20.5, 6.0])        # SURVEY 8c hand-checkable example
    vol = np.stack([m, np.zeros_like(m), m])
    a, b = get_ip_from_rvip_mask_3d(vol, keepdim=True)
    assert a == [[6.0, 11.0], None, [6.0, 11.0]] and b[1] is None
    a, b = get_ip_from_rvip_mask_3d(vol)
    assert len(a) == 2


def test_landmark_metrics_match_oracle_and_reference_golden(golden_dir):
    """Per-volume metrics kernel (rvip_landmark_metrics) against the oracle restatement and, through it, the golden
    vectors of the reference's own get_angle2x / get_distances / ... (float64: 1e-12)."""
    from cmr_landmark_detection_b200.extract import landmark_metrics_device
    from oracle import metrics_ref as M
    g = np.load(os.path.join(golden_dir, 'metrics_golden.npz'))
    for c in [str(c) for c in g['cases']]:
        gt = np.stack([g[c + '_gt_ant'], g[c + '_gt_inf']], axis=1)
        pr = np.stack([g[c + '_pr_ant'], g[c + '_pr_inf']], axis=1)
        spacing, thr, dim = [float(v) for v in g[c + '_params']]
        r = landmark_metrics_device(torch.from_numpy(gt).cuda(), torch.from_numpy(pr).cuda(), spacing, thr, dim)
        r = {k: v.cpu().numpy() for k, v in r.items()}

        def eq(a, b):
            return np.allclose(a, b, rtol=1e-12, atol=1e-11, equal_nan=True)
        assert eq(r['angle_gt'], g[c + '_angle_gt']) and eq(r['angle_pred'], g[c + '_angle_pr']), c
        for i, lm in enumerate(('ant', 'inf')):
            assert eq(r['dist'][i], g['%s_dist_%s' % (c, lm)]), c
            assert eq(r['dist_thr'][i], g['%s_dist_thr_%s' % (c, lm)]), c
            assert eq(r['dist_ub'][i], g['%s_ub_%s' % (c, lm)]), c
            tpr, ppv, tp, fn, fp = M.tpr_ppv(gt[:, i], pr[:, i], thr, spacing)
            assert abs(r['tpr'][i] - g[c + '_tpr'][i]) < 1e-15 and abs(r['ppv'][i] - g[c + '_ppv'][i]) < 1e-15, c
            assert list(r['counters'][i]) == [tp, fn, fp], c
        assert eq(r['mean_ip'][0, 0], g[c + '_mean_gt_ant']) and eq(r['mean_ip'][0, 1], g[c + '_mean_gt_inf']), c
        assert eq(r['mean_ip'][1, 0], g[c + '_mean_pr_ant']) and eq(r['mean_ip'][1, 1], g[c + '_mean_pr_inf']), c


def test_cc_filter_matches_oracle_and_reference_golden(golden_dir):
    """Largest-connected-component filter (rvip_cc_filter) against the oracle (bit-exact, both connectivities) and the
    golden outputs of the reference's own clean_3d_prediction_2d_cc (slices with a tied maximum: same area only)."""
    from cmr_landmark_detection_b200.data.Postprocess import clean_3d_prediction_2d_cc
    from cmr_landmark_detection_b200.extract import cc_filter_device
    from oracle import cc_ref
    g = np.load(os.path.join(golden_dir, 'cc_golden.npz'))
    for c in [str(c) for c in g['cases']]:
        vol, ref = g[c + '_in'], g[c + '_out']
        got = clean_3d_prediction_2d_cc(vol)
        assert got.shape == vol.shape and got.dtype == vol.dtype
        mine, ties = cc_ref.clean_2d_cc(vol, return_ties=True)
        assert np.array_equal(got, mine), c                                   # oracle: bit-exact incl. the tie rule
        for z in range(len(vol)):
            if ties[z]:     # tied maximum: OpenCV's internal numbering decides in the reference; same area required
                for val in (1, 2):
                    assert (got[z] == val).sum() == (ref[z] == val).sum(), (c, z)
            else:
                assert np.array_equal(got[z], ref[z]), (c, z)
        four = cc_filter_device(torch.from_numpy(vol).cuda(), 4).cpu().numpy()
        assert np.array_equal(four, cc_ref.clean_2d_cc(vol, connectivity=4)), c
    # full-size property: idempotent, and never adds pixels
    rng = np.random.default_rng(0)
    big = (rng.random((16, 256, 256)) < 0.3).astype(np.uint8) * rng.integers(1, 3, size=(16, 256, 256)).astype(np.uint8)
    d = torch.from_numpy(big).cuda()
    once = cc_filter_device(d)
    assert torch.equal(cc_filter_device(once), once)
    assert bool(((once == d) | (once == 0)).all())
