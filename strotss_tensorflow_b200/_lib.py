"""ctypes binding of libstrotss_b200.so (include/strotss_b200.h).  Fails loudly: there is no
fallback implementation anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libstrotss_b200.so")

NUM_SCALARS = 16
S_TOTAL, S_LOSS_C, S_LOSS_S, S_L_M, S_L_REMD, S_L_PALETTE = 0, 1, 2, 3, 4, 5
S_REMD_RX, S_REMD_RY, S_L_COV, S_L_MEAN, S_PAL_RX, S_PAL_RY, S_REMD_BRANCH, S_PAL_BRANCH = 6, 7, 8, 9, 10, 11, 12, 13

ERR_ARG, ERR_DISTANCE, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED = -1, -2, -3, -4, -5
DIST_CODES = {"cosine": 0, "l2": 1, "both": 2}       # keys of dist_metrics, nn/losses.py:27-28

_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> (restype, argtypes): must list every symbol declared in include/strotss_b200.h
SIGNATURES = {
    "strotss_create": (_i, [_i, C.POINTER(_vp)]),
    "strotss_destroy": (None, [_vp]),
    "strotss_last_error": (C.c_char_p, [_vp]),
    "strotss_version": (C.c_char_p, []),
    "strotss_workspace_bytes": (C.c_size_t, [_vp]),
    "strotss_launch_count": (_ll, [_vp]),
    "strotss_device_alloc": (_i, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "strotss_device_free": (_i, [_vp, _vp]),
    "strotss_profile_enable": (_i, [_vp, _i]),
    "strotss_profile_num_phases": (_i, []),
    "strotss_profile_phase_name": (C.c_char_p, [_i]),
    "strotss_profile_read": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(_ll)]),
    "strotss_comm_unique_id": (_i, [C.c_char_p]),
    "strotss_comm_init": (_i, [_vp, _i, _i, C.c_char_p]),
    "strotss_shard_rows": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i)]),
    "strotss_comm_transport": (_i, [_vp]),
    "strotss_set_style_target": (_i, [_vp, _vp, _i, _i, _ll, _vp]),
    "strotss_eval": (_i, [_vp, _vp, _ll, _vp, _ll, _i, _f, _vp, _vp, _ll, _vp, _vp, _vp]),
    "strotss_set_style_targets_grouped": (_i, [_vp, _vp, _ll, C.POINTER(_i), _i, _i, _vp]),
    "strotss_eval_grouped": (_i, [_vp, _vp, _ll, _vp, _ll, C.POINTER(_i), _i, _f, _vp, _vp, _vp, _ll, _vp]),
    "strotss_eval_host": (_i, [_vp, _vp, _vp, _i, _f, _vp, _vp, _vp]),
    "strotss_eval_host_submit": (_i, [_vp, _vp, _vp, _i, _f, _vp, _vp, C.POINTER(_ll)]),
    "strotss_eval_host_wait": (_i, [_vp, _ll]),
    "strotss_style_loss": (_i, [_vp, _vp, _ll, _i, _f, _vp, _vp, _ll, _vp]),
    "strotss_relaxed_emd": (_i, [_vp, _vp, _ll, _i, _vp, _ll, _i, _i, _i, _vp, _vp, _ll, _vp, _vp, _vp]),
    "strotss_moment_matching": (_i, [_vp, _vp, _ll, _i, _vp, _ll, _i, _i, _vp, _vp, _ll, _vp]),
    "strotss_self_similarity": (_i, [_vp, _vp, _ll, _vp, _ll, _i, _i, _vp, _vp, _ll, _vp]),
    "strotss_convert_rgb_to_yuv": (_i, [_vp, _vp, _ll, _i, _vp, _vp]),
    "strotss_sample": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _vp, _i, _i, _vp, _ll, _vp]),
    "strotss_sample_backward": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _vp, _i, _i, _vp, _ll, _vp]),
    "strotss_resize_bilinear": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _vp]),
    "strotss_make_laplacian": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "strotss_pyramid_fold": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i, _vp, _vp]),
    "strotss_pyramid_fold_backward": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i), _i, _vp, C.POINTER(_vp), _vp]),
    "strotss_rmsprop_step": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_ll), _f, _f, _f, _vp]),
    "strotss_debug_gemm": (_i, [_vp, _vp, _i, _vp, _i, _i, _f, _vp, _i, _vp]),
    "strotss_debug_gemm_ta": (_i, [_vp, _vp, _i, _vp, _i, _i, _f, _vp, _i, _vp]),
    "strotss_debug_tile_walk": (_i, [_i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), _i]),
    "strotss_debug_couples_pay": (_i, [_i, _i, _i, _i, _i]),
    "strotss_debug_ss_jobs": (_i, [_i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "strotss_debug_ss_copies": (_i, [_i, _i, _i, _i, C.POINTER(C.c_longlong), _i, C.POINTER(C.c_longlong)]),
}

_lib = None


def load():
    """dlopen the in-tree library and bind every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m strotss_tensorflow_b200.build` "
            "(nvcc, sm_100a).  This package has no CPU or eager fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class StrotssError(RuntimeError):
    pass


def check(lib, handle, code: int, what: str):
    if code == 0:
        return
    msg = lib.strotss_last_error(handle)
    text = msg.decode() if msg else ""
    if code == ERR_DISTANCE:
        raise KeyError(text or what)     # dist_metrics[distance] raises KeyError in the reference (nn/losses.py:74)
    if code == ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {text}")
    if code == ERR_ARG:
        raise ValueError(f"{what}: {text}")
    raise StrotssError(f"{what} failed (code {code}): {text}")
