"""ctypes binding of libwat.so (include/wat.h).  There is no fallback: if the library is missing or a
compute call is made without a CUDA device, an error is raised."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libwat.so")

WAT_FP32, WAT_BF16 = 0, 1
WAT_ERR_INVALID, WAT_ERR_CUDA, WAT_ERR_STATE, WAT_ERR_NOMEM = -1, -2, -3, -4


class WatConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_mels", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer", "at_low_compute", "at_dim",
        "n_class", "precision", "max_batch")]


class WatHeadConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "rep_dim", "n_layer", "inter_dim", "n_class", "mode", "n_time_head", "n_layer_head", "precision", "max_batch")]


# enum wat_head_mode (include/wat.h)
HEAD_MODES = {"lw_tr": 0, "lw_down_tr": 1, "mean_mlp": 2, "last_mlp": 3, "wa_mlp": 4, "mean_tr": 5, "last_tr": 6,
              "wa_tr": 7, "wa_down_tr": 8}


class WatError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libwat error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib: Optional[C.CDLL] = None

_vp, _i32, _i64, _fp = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
_SIGS = {
    "wat_abi_version": (C.c_int, []),
    "wat_last_error": (C.c_char_p, []),
    "wat_create": (C.c_int, [C.POINTER(WatConfig), C.POINTER(_vp)]),
    "wat_set_weight": (C.c_int, [_vp, C.c_char_p, _fp, _i64]),
    "wat_finalize": (C.c_int, [_vp]),
    "wat_destroy": (C.c_int, [_vp]),
    "wat_logmel": (C.c_int, [_vp, _fp, _i64, _vp, _i32, _i32, _i32, _i32, _i32, _fp, _vp]),
    "wat_encoder": (C.c_int, [_vp, _fp, _i32, _fp, _fp, _vp]),
    "wat_tltr": (C.c_int, [_vp, _fp, _i32, _i32, _i32, _i32, _i32, _fp, _vp]),
    "wat_head_create": (C.c_int, [C.POINTER(WatHeadConfig), C.POINTER(_vp)]),
    "wat_head_forward": (C.c_int, [_vp, _fp, _i32, _i32, _fp, _vp]),
    "wat_tag": (C.c_int, [_vp, _fp, _i64, _vp, _i32, _i32, _i32, _fp, _vp]),
    "wat_tag_host": (C.c_int, [_vp, _fp, _i64, _vp, _i32, _i32, _i32, _fp]),
    "wat_tag_pcm16": (C.c_int, [_vp, _fp, _i64, _vp, _i32, _i32, _i32, _fp, _vp]),
    "wat_tag_host_pcm16": (C.c_int, [_vp, _fp, _i64, _vp, _i32, _i32, _i32, _fp]),
    "wat_tag_host_submit": (C.c_int, [_vp, _fp, _i64, _vp, _i32, _i32, _i32, _fp, _vp]),
    "wat_tag_host_submit_pcm16": (C.c_int, [_vp, _fp, _i64, _vp, _i32, _i32, _i32, _fp, _vp]),
    "wat_tag_host_wait": (C.c_int, [_vp, _i64]),
    "wat_workspace_bytes": (_i64, [_vp]),
    "wat_kernel_launches": (_i64, [_vp]),
    "wat_num_sms": (C.c_int, [_vp]),
    "wat_profile": (C.c_int, [_vp, _i32]),
    "wat_profile_classes": (C.c_int, []),
    "wat_profile_class_name": (C.c_char_p, [_i32]),
    "wat_profile_read": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "wat_dbg_gemm": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "wat_dbg_gemm_bf16": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "wat_dbg_ln_slices": (C.c_int, [_i32, _i32, _i32, _i32]),
    "wat_dbg_ln_gemm": (C.c_int, [_fp] * 13 + [_i32] * 7 + [_vp]),
    "wat_dbg_attention": (C.c_int, [_fp, _fp, _fp, _fp, _i32, _i32, _i32, _i32, _vp]),
    "wat_dbg_attention_repeats": (C.c_int, []),
    "wat_dbg_tma_overlap_probe": (C.c_int, []),
}
EXPORTS = tuple(_SIGS)


def lib() -> C.CDLL:
    """Load libwat.so (once).  Raises if it has not been built: run `make -C whisper-at_b200`
    or `python -c "import __graft_entry__ as g; g.build()"`."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA library of whisper_at (B200) is not built and there is no CPU "
                f"fallback. Build it with `make -C whisper-at_b200`.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise WatError(rc, lib().wat_last_error().decode("utf-8", "replace"))
