"""ctypes binding of libeabnet_b200.so (the C ABI in include/eabnet_b200.h).  No torch types cross it."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libeabnet_b200.so")


class EabConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "k1_t", "k1_f", "k2_t", "k2_f", "c", "M", "embed_dim", "kd1", "cd1", "d_feat", "p", "q", "is_causal",
        "is_u2", "bf_type", "topo_type", "intra_connect", "norm_type", "n_freq")]


class EabGagConfig(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("cin", "k1_t", "k1_f", "k2_t", "k2_f", "c", "kd1", "cd1", "d_feat", "p", "q",
                                        "n_dilas")] + [("dilas", C.c_int * 8)] +
                [(n, C.c_int) for n in ("fft_num", "is_u2", "is_causal", "is_squeezed", "acti_type", "intra_connect",
                                        "norm_type")])


# every symbol include/eabnet_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_F = C.c_void_p          # float* passed as raw address
SYMBOLS = {
    "eab_create": (C.c_int, [C.POINTER(EabConfig), C.POINTER(_P)]),
    "eab_destroy": (None, [_P]),
    "eab_param_count": (C.c_int, [_P]),
    "eab_param_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int64 * 4),
                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "eab_set_param": (C.c_int, [_P, C.c_char_p, _F, C.c_int64]),
    "eab_commit_params": (C.c_int, [_P, _P]),
    "eab_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int]),
    "eab_forward": (C.c_int, [_P, _F, _F, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "eab_stft": (C.c_int, [_F, _F, C.c_int, C.c_int, C.c_int, _P]),
    "eab_istft": (C.c_int, [_F, _F, C.c_int, C.c_int, _P]),
    "eab_enhance_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int]),
    "eab_enhance": (C.c_int, [_P, _F, _F, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "eab_enhance_host": (C.c_int, [_P, _F, _F, C.c_int, C.c_int, _P]),
    "eab_enhance_host_batches": (C.c_int, [_P, C.POINTER(_F), C.POINTER(_F), C.c_int, C.c_int, C.c_int, _P]),
    "eab_stream_state_bytes": (C.c_size_t, [_P, C.c_int]),
    "eab_stream_reset": (C.c_int, [_P, _P, C.c_size_t, C.c_int, _P]),
    "eab_stream_step": (C.c_int, [_P, _P, C.c_size_t, _F, _F, C.c_int, _P]),
    "eab_stream_step_spec": (C.c_int, [_P, _P, C.c_size_t, _F, _F, C.c_int, _P]),
    "eab_stream_step_pcm16": (C.c_int, [_P, _P, C.c_size_t, _P, _P, C.c_int, _P]),
    "eab_stream_reset_one": (C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_int, _P]),
    "eab_gag_create": (C.c_int, [C.POINTER(EabGagConfig), C.POINTER(_P)]),
    "eab_gag_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int]),
    "eab_gag_forward": (C.c_int, [_P, _F, C.POINTER(C.c_int64 * 4), _F, _F, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "eab_enhance_postnet_workspace_bytes": (C.c_size_t, [_P, _P, C.c_int, C.c_int]),
    "eab_enhance_postnet": (C.c_int, [_P, _P, C.c_int, _F, _F, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "eab_enhance_host_pcm16": (C.c_int, [_P, _P, C.c_int, _P, C.POINTER(C.c_int), _P, C.c_int, C.c_int, _P]),
    "eab_enhance_host_batches_pcm16": (C.c_int, [_P, C.POINTER(_F), C.POINTER(C.c_int), C.POINTER(_F), C.c_int, C.c_int, C.c_int, _P]),
    "eab_head_forward": (C.c_int, [_F, _F, _F, _F, _F, _F, _F, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "eab_head_backward_workspace_bytes": (C.c_size_t, [C.c_int]),
    "eab_head_backward": (C.c_int, [_F, _F, _F, _F, _F, _F, _F, _F, _F, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "eab_loss_com_mag_mse": (C.c_int, [_F, _F, _P, C.c_int64, C.c_int, C.c_int, C.c_int, _F, _P, _P]),
    "eab_loss_com_mag_mse_backward": (C.c_int, [_F, _F, _P, C.c_int64, C.c_int, C.c_int, C.c_int, _F, _F, _P]),
    "eab_loss_com_mag_mse_fm": (C.c_int, [_F, _F, _P, C.c_int64, C.c_int, C.c_int, C.c_int, _F, _P, _P]),
    "eab_loss_com_mag_mse_fm_backward": (C.c_int, [_F, _F, _P, C.c_int64, C.c_int, C.c_int, C.c_int, _F, _F, _P]),
    "eab_wav_info": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int),
                              C.POINTER(C.c_int)]),
    "eab_wav_decode": (C.c_int, [_P, C.c_size_t, _F, _P]),
    "eab_wav_encode_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "eab_wav_encode": (C.c_int, [_F, _P, C.c_int64, C.c_int, C.c_int, _P, C.c_size_t]),
    "eab_resample_length": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "eab_resample": (C.c_int, [_F, _F, C.c_int, C.c_int64, C.c_int, C.c_int, _P]),
    "eab_gag_stream_step_spec": (C.c_int, [_P, _P, C.c_size_t, _F, _F, _F, C.c_int, _P]),
    "eab_stream_step_postnet": (C.c_int, [_P, _P, C.c_size_t, _P, _P, C.c_size_t, C.c_int, _F, _F, C.c_int, _P]),
    "eab_stream_step_postnet_pcm16": (C.c_int, [_P, _P, C.c_size_t, _P, _P, C.c_size_t, C.c_int, _P, _P, C.c_int, _P]),
    "eab_norm_stats_count": (C.c_int, [_P]),
    "eab_norm_stats": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int64), _P, _P]),
    "eab_last_launch_count": (C.c_int, [_P]),
    "eab_debug_tap": (C.c_int64, [_P, C.c_char_p, _F, C.c_int64, _P]),
    "eab_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "eab_debug_counters": (C.c_int, [_P, C.POINTER(C.c_uint64 * 16)]),
    "eab_profile_enable": (C.c_int, [_P, C.c_int]),
    "eab_profile_summary": (C.c_int64, [_P, C.c_char_p, C.c_int64]),
    "eab_last_error": (C.c_char_p, []),
    "eab_build_info": (C.c_char_p, []),
}

_lib = None


def load():
    """Load the CUDA library; fail loudly if it has not been built (there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "eabnet_b200: %s is missing - build it with `python -m eabnet_b200.build` "
                "(or __graft_entry__.build()); there is no CPU / PyTorch fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().eab_last_error()
        raise RuntimeError("eabnet_b200%s: %s" % (" (" + what + ")" if what else "", msg.decode() if msg else "error"))
