"""ctypes binding of librvip_b200.so (include/rvip.h).  No torch types cross this boundary: only raw
device pointers, sizes and a cudaStream_t.  There is deliberately NO fallback: if the CUDA library is
missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.dirname(_HERE)
LIB_PATH = os.path.join(PKG_DIR, 'librvip_b200.so')
CSRC_DIR = os.path.join(PKG_DIR, 'csrc')

RVIP_MAX_DEPTH = 8
NUM_KERNEL_CLASSES = 10
LOSS_KINDS = {'mse': 0, 'masked': 1, 'weighted': 2, 'bce_dice': 3}


class RvipError(RuntimeError):
    pass


class rvip_cfg(C.Structure):
    _fields_ = [('H', C.c_int), ('W', C.c_int), ('in_ch', C.c_int), ('classes', C.c_int), ('depth', C.c_int),
                ('filters', C.c_int), ('batch_norm', C.c_int), ('bn_first', C.c_int), ('use_upsample', C.c_int),
                ('precision', C.c_int), ('dropout', C.c_float * RVIP_MAX_DEPTH), ('dropout_mid', C.c_float),
                ('bn_momentum', C.c_float), ('bn_eps', C.c_float)]


def build_library(force: bool = False) -> str:
    """Compiles csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(['make', '-C', CSRC_DIR, 'clean'])
    subprocess.check_call(['make', '-C', CSRC_DIR, '-j8'])
    return LIB_PATH


_lib = None

_VP, _I, _LL, _F, _SZ = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
_SIGS = {
    'rvip_last_error': (C.c_char_p, []),
    'rvip_abi_version': (_I, []),
    'rvip_create': (_I, [C.POINTER(rvip_cfg), C.POINTER(_VP)]),
    'rvip_destroy': (None, [_VP]),
    'rvip_param_count': (_LL, [_VP]),
    'rvip_state_count': (_LL, [_VP]),
    'rvip_num_tensors': (_I, [_VP]),
    'rvip_tensor_info': (_I, [_VP, _I, C.c_char_p, _I, C.POINTER(_I), C.POINTER(_LL), C.POINTER(_I),
                              C.POINTER(_I * 4)]),
    'rvip_workspace_bytes': (_SZ, [_VP, _I, _I]),
    'rvip_bind': (_I, [_VP, _VP, _VP, _VP, _VP, _SZ, _I, _I]),
    'rvip_pack_weights': (_I, [_VP, _VP]),
    'rvip_predict': (_I, [_VP, _VP, _VP, _VP]),
    'rvip_train_step': (_I, [_VP, _VP, _VP, _VP, _I, _F, C.c_uint64, _VP, _VP, _VP]),
    'rvip_set_loss_weights': (_I, [_VP, _F, _F]),
    'rvip_heat_stats': (_I, [_VP, _VP, _VP, _LL, _I, _I, _I, _F, _VP, _VP]),
    'rvip_adam_step': (_I, [_VP, _VP, _VP, _F, _F, _F, _F, _LL, _F, _VP]),
    'rvip_set_inline_adam': (_I, [_VP, _VP, _VP, _F, _F, _F, _F, _LL, _F]),
    'rvip_adam_bucket': (_I, [_VP, _I, _VP, _VP, _F, _F, _F, _F, _LL, _F, _VP]),
    'rvip_sgd_step': (_I, [_VP, _VP, _F, _F, _I, _F, _VP]),
    'rvip_num_buckets': (_I, [_VP]),
    'rvip_bucket': (_I, [_VP, _I, C.POINTER(_LL), C.POINTER(_LL)]),
    'rvip_set_bucket_event': (_I, [_VP, _I, _VP]),
    'rvip_extract_scratch_bytes': (_SZ, [_I, _I]),
    'rvip_extract': (_I, [_VP, _I, _I, _I, _I, _F, _VP, _VP, _VP, _VP, _VP, _VP]),
    'rvip_label_map': (_I, [_VP, _LL, _I, _F, _VP, _VP]),
    'rvip_cc_scratch_bytes': (_SZ, [_I, _I, _I]),
    'rvip_cc_filter': (_I, [_VP, _I, _I, _I, _I, _VP, _VP, _VP]),
    'rvip_landmark_metrics': (_I, [_VP, _VP, _I, C.c_double, C.c_double, C.c_double, _VP, _VP, _VP, _VP, _VP, _VP]),
    'rvip_debug_buffer': (_I, [_VP, C.c_char_p, _I, C.POINTER(_VP), C.POINTER(_LL), C.POINTER(_I)]),
    'rvip_dropout_mask': (_I, [C.c_uint64, C.c_uint32, _F, _LL, _VP, _VP]),
    'rvip_dropout_site': (_I, [_VP, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(_F)]),
    'rvip_profile': (_I, [_VP, _I]),
    'rvip_profile_read': (_I, [_VP, C.POINTER(_F * NUM_KERNEL_CLASSES), C.POINTER(_LL * NUM_KERNEL_CLASSES)]),
    'rvip_profile_detail': (C.c_char_p, [_VP]),
    'rvip_kernel_class_name': (C.c_char_p, [_I]),
    'rvip_launch_count': (_LL, [_VP]),
    'rvip_conv3x3_tc': (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _I, _VP, _I, _I, _I, _I, _I, _VP]),
    'rvip_conv3x3_halo': (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _I, _VP, _I, _I, _I, _I, _I, _VP]),
    'rvip_conv3x3_halo_debug': (_I, [_VP]),
    'rvip_conv3x3_row': (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _I, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    'rvip_wgrad3x3_row': (_I, [_VP, _VP, _I, _I, _VP, _VP, _I, _I, _I, _I, _VP]),
    'rvip_wgrad3x3_tc': (_I, [_VP, _VP, _I, _I, _VP, _VP, _I, _I, _I, _I, _VP]),
    'rvip_wgrad3x3_halo': (_I, [_VP, _VP, _I, _I, _VP, _VP, _I, _I, _I, _I, _VP]),
    'rvip_upconv_wgrad_halo': (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    'rvip_upconv3x3_halo': (_I, [_I, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


def lib():
    """Loads the CUDA library (once). Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RvipError('%s is missing: build it with __graft_entry__.build() / make -C %s. '
                            'This package has no CPU fallback.' % (LIB_PATH, CSRC_DIR))
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        raise RvipError(lib().rvip_last_error().decode('utf-8', 'replace'))


def ptr(t):
    """torch tensor / None -> device (or host) pointer as int."""
    return None if t is None else C.c_void_p(t.data_ptr())
