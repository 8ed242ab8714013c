"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol include/rvip.h declares,
builds the network plan without a GPU, and the host mirror refuses to run without CUDA (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from cmr_landmark_detection_b200.runtime import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, 'include', 'rvip.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(rvip_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 28
    lib = C.CDLL(ffi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), 'librvip_b200.so does not export %s' % name
    assert set(ffi.EXPORTED_SYMBOLS) == declared, (set(ffi.EXPORTED_SYMBOLS) ^ declared)


def test_plan_and_tensor_table_without_gpu():
    L = ffi.lib()
    cfg = ffi.rvip_cfg(H=128, W=128, in_ch=1, classes=2, depth=4, filters=32, batch_norm=1, bn_first=0, use_upsample=1,
                       precision=1, dropout_mid=0.5, bn_momentum=0.99, bn_eps=1e-3)
    h = C.c_void_p()
    ffi.check(L.rvip_create(C.byref(cfg), C.byref(h)))
    assert L.rvip_param_count(h) == 8635842 and L.rvip_state_count(h) == 5888      # reference model.summary()
    assert L.rvip_num_tensors(h) == 118
    names = []
    for i in range(118):
        nm = C.create_string_buffer(128)
        st, off, nd, dims = C.c_int(), C.c_longlong(), C.c_int(), (C.c_int * 4)()
        ffi.check(L.rvip_tensor_info(h, i, nm, 128, C.byref(st), C.byref(off), C.byref(nd), C.byref(dims)))
        names.append(nm.value.decode())
    assert names[0] == 'enc0.conv_a/kernel' and names[-1] == 'head/bias' and names[2] == 'enc0.conv_a/bn/gamma'
    # buckets tile the gradient buffer back to front
    end = 8635842
    for i in range(L.rvip_num_buckets(h)):
        o, c = C.c_longlong(), C.c_longlong()
        ffi.check(L.rvip_bucket(h, i, C.byref(o), C.byref(c)))
        assert o.value + c.value == end
        end = o.value
    assert end == 0
    assert L.rvip_workspace_bytes(h, 32, 1) > L.rvip_workspace_bytes(h, 32, 0) > 0
    L.rvip_destroy(h)


def test_unsupported_configs_fail_loudly():
    L = ffi.lib()
    for kw in (dict(use_upsample=0, filters=96), dict(H=100), dict(filters=48),
               dict(classes=9)):
        base = dict(H=128, W=128, in_ch=1, classes=2, depth=4, filters=32, batch_norm=1, bn_first=0, use_upsample=1,
                    precision=1, dropout_mid=0.5, bn_momentum=0.99, bn_eps=1e-3)
        base.update(kw)
        h = C.c_void_p()
        assert L.rvip_create(C.byref(ffi.rvip_cfg(**base)), C.byref(h)) != 0
        assert len(L.rvip_last_error()) > 0


def test_conv2d_transpose_plan():
    """USE_UPSAMPLE falsy (Conv2DTranspose decoder): planned in bf16 mode when every up-conv fits the phase-decomposed
    kernels, with Keras' (kh, kw, out, in) kernel shape in the tensor table; fp32 mode plans it on the CUDA-core kernels."""
    L = ffi.lib()
    base = dict(H=256, W=256, in_ch=1, classes=2, depth=4, filters=32, batch_norm=1, bn_first=0, use_upsample=0,
                precision=1, dropout_mid=0.5, bn_momentum=0.99, bn_eps=1e-3)
    h = C.c_void_p()
    ffi.check(L.rvip_create(C.byref(ffi.rvip_cfg(**base)), C.byref(h)))
    assert L.rvip_param_count(h) == 8635842
    found = {}
    for i in range(L.rvip_num_tensors(h)):
        name = C.create_string_buffer(128)
        st, off, nd = C.c_int(), C.c_longlong(), C.c_int()
        dims = (C.c_int * 4)()
        ffi.check(L.rvip_tensor_info(h, i, name, 128, C.byref(st), C.byref(off), C.byref(nd), dims))
        found[name.value.decode()] = tuple(dims[:nd.value])
    assert found['dec0.upconv/kernel'] == (3, 3, 256, 512) and found['dec3.upconv/kernel'] == (3, 3, 32, 64)
    assert found['dec0.conv_a/kernel'] == (3, 3, 512, 256)
    L.rvip_destroy(h)
    # fp32 parity mode plans the same tensor table (CUDA-core convolution over the zero-stuffed input)
    h = C.c_void_p()
    ffi.check(L.rvip_create(C.byref(ffi.rvip_cfg(**dict(base, precision=0))), C.byref(h)))
    assert L.rvip_param_count(h) == 8635842
    L.rvip_destroy(h)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from cmr_landmark_detection_b200.models.Unets import create_unet
    with pytest.raises(ffi.RvipError):
        create_unet({'DIM': [32, 32], 'DEPTH': 2, 'FILTERS': 32, 'BATCH_NORMALISATION': True, 'ACTIVATION': 'relu',
                     'MASK_CLASSES': 2})


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'cmr_landmark_detection_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dp, f)).read()
                assert 'oracle' not in src.replace('the oracle', ''), os.path.join(dp, f)


def test_deferred_weight_gradient_plan(monkeypatch):
    """RVIP_DEFER_WGRAD (opt-in backward schedule): the named deep-level layers get dz buffers of their own, so the
    training workspace grows by exactly those tensors (rounded to the carver's 1 KB granules); inference is untouched."""
    L = ffi.lib()
    base = dict(H=256, W=256, in_ch=1, classes=2, depth=4, filters=32, batch_norm=1, bn_first=0, use_upsample=1,
                precision=1, dropout_mid=0.5, bn_momentum=0.99, bn_eps=1e-3)

    def sizes():
        h = C.c_void_p()
        ffi.check(L.rvip_create(C.byref(ffi.rvip_cfg(**base)), C.byref(h)))
        out = (L.rvip_workspace_bytes(h, 32, 1), L.rvip_workspace_bytes(h, 32, 0))
        L.rvip_destroy(h)
        return out

    monkeypatch.delenv('RVIP_DEFER_WGRAD', raising=False)
    t0, i0 = sizes()
    monkeypatch.setenv('RVIP_DEFER_WGRAD', 'mid.conv_b,dec0.conv_a')
    t1, i1 = sizes()
    extra = 32 * 16 * 16 * 512 * 2 + 32 * 32 * 32 * 256 * 2          # bf16 dz of mid.conv_b and dec0.conv_a at batch 32
    assert i1 == i0 and extra <= t1 - t0 <= extra + 4096
    monkeypatch.setenv('RVIP_DEFER_WGRAD', 'none')
    assert sizes() == (t0, i0)
