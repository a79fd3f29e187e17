"""ctypes binding of libssr_b200.so (the C ABI declared in include/ssr_b200.h).

There is no fallback: if the library is missing it is built with nvcc; if that fails, importing raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

c_i32, c_i64, c_f32p, c_vp, c_cp = C.c_int32, C.c_int64, C.POINTER(C.c_float), C.c_void_p, C.c_char_p
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)

SSR_WAVLM, SSR_WHISPER_ENC = 0, 1
SSR_FEAT_NORM_GROUP, SSR_FEAT_NORM_LAYER = 0, 1


class ModelDesc(C.Structure):
    _fields_ = [
        ("family", c_i32), ("hidden", c_i32), ("layers", c_i32), ("heads", c_i32), ("ffn", c_i32),
        ("feat_norm", c_i32), ("stable_ln", c_i32), ("do_normalize", c_i32), ("n_mels", c_i32),
        ("reserved", c_i32 * 7),
    ]


class Weight(C.Structure):
    _fields_ = [("name", c_cp), ("data", c_vp), ("numel", c_i64)]


SSR_AUG_NONE, SSR_AUG_SPEED, SSR_AUG_NOISE, SSR_AUG_VOLUME, SSR_AUG_PITCH = 0, 1, 2, 3, 4


class AugOp(C.Structure):
    _fields_ = [("kind", c_i32), ("new_rate", c_i32), ("factor", C.c_float), ("reserved", c_i32),
                ("seed", C.c_uint64)]


# name -> (restype, argtypes); must list every symbol include/ssr_b200.h declares (tests check this).
SIGNATURES = {
    "ssr_create": (c_i32, [C.POINTER(ModelDesc), C.POINTER(Weight), c_i32, c_i32, C.POINTER(c_vp)]),
    "ssr_destroy": (None, [c_vp]),
    "ssr_last_error": (c_cp, [c_vp]),
    "ssr_set_option": (c_i32, [c_vp, c_cp, c_i32]),
    "ssr_wavlm_pooled": (c_i32, [c_vp, c_vp, c_i64, c_i32p, c_i32, c_vp, c_vp]),
    "ssr_whisper_enc_pooled": (c_i32, [c_vp, c_vp, c_i64, c_i32p, c_i32, c_vp, c_vp]),
    "ssr_logmel": (c_i32, [c_vp, c_vp, c_i64, c_i32p, c_i32, c_vp, c_vp]),
    "ssr_wavlm_pooled_host": (c_i32, [c_vp, c_vp, c_i64, c_i32p, c_i32, c_vp]),
    "ssr_whisper_enc_pooled_host": (c_i32, [c_vp, c_vp, c_i64, c_i32p, c_i32, c_vp]),
    "ssr_decoder_layers": (c_i32, [c_vp]),
    "ssr_whisper_full": (c_i32, [c_vp, c_vp, c_i64, c_i32p, c_i32, c_vp, c_vp, c_vp]),
    "ssr_whisper_full_host": (c_i32, [c_vp, c_vp, c_i64, c_i32p, c_i32, c_vp, c_vp]),
    "ssr_num_frames": (c_i32, [c_vp, c_i32]),
    "ssr_launch_count": (c_i64, [c_vp]),
    "ssr_wavlm_rel_bucket": (c_i32, [c_i32]),
    "ssr_tuning_set": (c_i32, [c_cp, c_i32]),
    "ssr_profile_fetch": (c_cp, [c_vp]),
    "ssr_gemm_bf16": (c_i32, [c_i32, c_vp, c_i64, c_i64, c_vp, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp,
                              c_i32, c_vp, c_cp, c_i32]),
    "ssr_gemm_bf16_pool": (c_i32, [c_i32, c_vp, c_i64, c_vp, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp, c_i32,
                                   c_vp, c_i32, c_vp, c_vp, c_i64, c_i32, c_vp, c_cp, c_i32]),
    "ssr_layernorm": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_cp, c_i32]),
    "ssr_attention": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_cp,
                              c_i32]),
    "ssr_pool_mean": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_i64, c_vp, c_cp, c_i32]),
    "ssr_debug_fetch": (c_i64, [c_vp, c_cp, c_vp, c_i64, c_i64p, c_i32p]),
    "ssr_head_param_count": (c_i64, [c_i32, c_i32, c_i32]),
    "ssr_head_work_bytes": (c_i64, [c_i64, c_i32, c_i32, c_i32]),
    "ssr_head_scaler_stats": (c_i32, [c_vp, c_i64, c_i32, c_i64, c_vp, c_vp, c_vp, c_cp, c_i32]),
    "ssr_head_grad": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64,
                              c_vp, c_cp, c_i32]),
    "ssr_head_adam": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                              c_i32, c_vp, c_cp, c_i32]),
    "ssr_head_predict": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp,
                                 c_cp, c_i32]),
    "ssr_resample_length": (c_i32, [c_i32, c_i32, c_i32]),
    "ssr_augment_out_length": (c_i32, [C.POINTER(AugOp), c_i32, c_i32]),
    "ssr_augment_work_bytes": (c_i64, [c_i32p, c_i32, C.POINTER(AugOp), c_i32]),
    "ssr_augment": (c_i32, [c_vp, c_i64, c_i32p, c_i32, C.POINTER(AugOp), c_i32, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64,
                            c_i32p, c_vp, c_cp, c_i32]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """Load (building first if needed) the CUDA library. Raises if it cannot be produced."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing and could not be built; ssr_b200 has no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == ABI drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
