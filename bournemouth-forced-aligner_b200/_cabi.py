"""ctypes binding of libbfa_b200.so (include/bfa_b200.h).  No fallback: if the CUDA library is
missing or cannot be loaded this module raises, it never routes to a CPU implementation."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

HERE = Path(__file__).resolve().parent
# BFA_B200_LIB selects another build of the same library (development: e.g. the -DBFA_PHASE_PROF build)
LIB_PATH = Path(os.environ["BFA_B200_LIB"]) if os.environ.get("BFA_B200_LIB") else HERE / "lib" / "libbfa_b200.so"

BFA_OK, BFA_E_INVALID, BFA_E_UNSUPPORTED, BFA_E_WORKSPACE, BFA_E_CUDA = 0, -1, -2, -3, -4
ST_OK, ST_EMPTY_TARGET, ST_TOO_SHORT, ST_PROPORTIONAL, ST_SEGMENTED = 0, 1, 2, 3, 4
ST_DEFERRED, ST_UNSUPPORTED = 5, 6
ST_DEGENERATE, ST_STAMP_OVERFLOW = 8, 16
MODE_FULL, MODE_SIMPLE = 0, 1
FLAG_EXACT_ONLY, HINT_NO_SIL = 1, 2
FLAG_UNFUSED_CONF, FLAG_NO_SPEC, FLAG_FILL_ONLY = 4, 8, 16
FLAG_NO_DIRECT, FLAG_DIRECT_ONLY, FLAG_PIPELINED = 32, 64, 128
MAX_C, MAX_L = 256, 8192


class BfaParams(C.Structure):
    _fields_ = [
        ("blank_id", C.c_int32), ("silence_id", C.c_int32), ("silence_anchors", C.c_int32),
        ("ignore_noise", C.c_int32), ("truly_forced", C.c_int32), ("boost_targets", C.c_int32),
        ("enforce_minimum", C.c_int32), ("max_blanks", C.c_int32),
        ("boost_factor", C.c_float), ("min_log_prob", C.c_float), ("neg_inf", C.c_float),
        ("sub_boost", C.c_float), ("boundary_pad", C.c_int32), ("min_speech_frames", C.c_int32),
        ("mode", C.c_int32), ("reserved", C.c_int32),
    ]


class BfaShape(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("max_T", C.c_int32), ("max_N", C.c_int32),
                ("total_frames", C.c_int64), ("max_stamps", C.c_int32), ("reserved", C.c_int32)]


class BfaError(RuntimeError):
    code = None        # the library's return code when the error came from a C-ABI call


_P = C.c_void_p
_SIGNATURES = {
    "bfa_version": (C.c_int, []),
    "bfa_strerror": (C.c_char_p, [C.c_int]),
    "bfa_last_cuda_error": (C.c_char_p, []),
    "bfa_host_last_error": (C.c_char_p, []),
    "bfa_sizeof_params": (C.c_int, []),
    "bfa_launch_count": (C.c_int64, []),
    "bfa_default_params": (None, [C.POINTER(BfaParams), C.c_int32, C.c_int32]),
    "bfa_workspace_bytes": (C.c_size_t, [C.POINTER(BfaParams), C.POINTER(BfaShape)]),
    "bfa_align_batch": (C.c_int, [C.POINTER(BfaParams), C.POINTER(BfaShape)] + [_P] * 13 + [_P, C.c_size_t, _P]),
    "bfa_align_batch_logits": (C.c_int, [C.POINTER(BfaParams), C.POINTER(BfaShape)] + [_P] * 14 + [_P, C.c_size_t, _P]),
    "bfa_viterbi_paths_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "bfa_viterbi_paths": (C.c_int, [C.POINTER(BfaParams), C.c_int32, C.c_int32, C.c_int32, C.c_int32] + [_P] * 13
                          + [_P, C.c_size_t, _P]),
    "bfa_confidence_batch": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.c_int32, _P, _P]),
    "bfa_confidence_batch_lse": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, _P]),
    "bfa_soft_boundaries_batch_lse": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "bfa_alignment_score_batch": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P]),
    "bfa_stitch_log_softmax": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64, _P, _P, C.c_int64, _P]),
    "bfa_assort_batch": (C.c_int, [C.POINTER(BfaParams), C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_int32, _P]),
    "bfa_align_batch_host": (C.c_int, [C.POINTER(BfaParams), C.POINTER(BfaShape)] + [_P] * 13 + [C.c_int32, C.c_int32]),
    "bfa_host_release": (None, []),
    "bfa_debug_phases": (C.c_int, [_P, C.c_int]),
    "bfa_debug_warps": (C.c_int, [_P, C.c_int]),
    "bfa_soft_boundaries_batch": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "bfa_profile_read_aux": (C.c_int, [_P]),
    "bfa_debug_item_counts": (C.c_int, [_P]),
    "bfa_debug_ctas": (C.c_int, [_P, C.c_int]),
    "bfa_debug_fin": (C.c_int, [_P, C.c_int]),
    "bfa_profile_enable": (None, [C.c_int]),
    "bfa_profile_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
}

_lib = None


def lib():
    """Load libbfa_b200.so (built in-tree by build.py / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise BfaError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (or bournemouth-forced-aligner_b200/build.py) "
                           "to compile the CUDA library; there is no CPU fallback")
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if l.bfa_sizeof_params() != C.sizeof(BfaParams):
            raise BfaError("BfaParams layout mismatch between _cabi.py and libbfa_b200.so")
        _lib = l
    return _lib


def exported_symbols():
    return list(_SIGNATURES)


def check(rc: int, host: bool = False):
    if rc == BFA_OK:
        return
    l = lib()
    msg = l.bfa_strerror(rc).decode()
    if rc == BFA_E_CUDA:
        detail = (l.bfa_host_last_error() if host else l.bfa_last_cuda_error()).decode()
        msg = f"{msg}: {detail}"
    err = BfaError(f"libbfa_b200: {msg} (code {rc})")
    err.code = rc
    raise err


def default_params(blank_id: int, silence_id) -> BfaParams:
    p = BfaParams()
    lib().bfa_default_params(C.byref(p), int(blank_id), -1 if silence_id is None else int(silence_id))
    return p
