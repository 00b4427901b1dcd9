"""ctypes binding of ``libcoral_b200.so`` (the C ABI declared in ``include/coral_b200.h``).

The product has no CPU path: if the CUDA library is missing this module raises, and every
public entry point of the package goes through it.
"""

from __future__ import annotations

import ctypes as C
import os

from ._build import LIB_PATH

_lib = None

OK, EARG, EIO, ECUDA, ECAP = 0, -1, -2, -3, -4

_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double

SIGNATURES = {
    "coral_last_error": (C.c_char_p, []),
    "coral_abi_version": (_i32, []),
    "coral_lm_load_arpa": (_i32, [C.c_char_p, _i32, C.POINTER(_vp)]),
    "coral_lm_load_kenlm_binary": (_i32, [C.c_char_p, _i32, C.POINTER(_vp)]),
    "coral_lm_load": (_i32, [C.c_char_p, _i32, C.POINTER(_vp)]),
    "coral_lm_free": (_i32, [_vp]),
    "coral_lm_info": (_i32, [_vp, C.POINTER(_i32), _vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "coral_lm_contains": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "coral_lm_score_sentences": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "coral_decoder_create": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _i32, C.POINTER(_vp)]),
    "coral_decoder_free": (_i32, [_vp]),
    "coral_decoder_set_params": (_i32, [_vp, _f64, _f64, _f64, _i32]),
    "coral_decoder_info": (_i32, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "coral_ctc_beam_decode": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f64, _f64, _i32, _i32, _i32,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp]),
    "coral_decoder_tokens_to_text": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "coral_host_pack_rows": (_i32, [_vp, _vp, _vp, _i64, _vp, _i32]),
    "coral_py_string_list": (C.py_object, [_vp, _i32, _vp, _i64]),
    "coral_py_logits_rows": (_i64, [C.py_object, _i32, _vp, _vp, _vp]),
    "coral_normaliser_create": (_i32, [_vp, _i64, _vp, _vp, _i64, _i32, _i32, C.POINTER(_vp)]),
    "coral_normaliser_free": (_i32, [_vp]),
    "coral_normaliser_run": (_i32, [_vp, _vp, _vp, _i64, _i32, C.POINTER(_i64)]),
    "coral_normaliser_fetch": (_i32, [_vp, _vp, _vp, _vp]),
    "coral_normaliser_unicode_version": (C.c_char_p, []),
    "coral_ctc_greedy": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "coral_ctc_collapse": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "coral_edit_counts": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i64, _i32, _vp, _vp, _vp]),
    "coral_edit_counts_spans": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i64, _i32, _vp, _vp, _vp]),
}


class CoralError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"coral_b200 error {code}: {message}")
        self.code = code
        self.message = message


def lib_path() -> str:
    return os.environ.get("CORAL_B200_LIB", LIB_PATH)


def load():
    """Load the CUDA library. Raises if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"coral_b200: the CUDA library {path} is missing. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` -- there is no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    # the entry points that call back into CPython must run with the GIL held
    pydll = C.PyDLL(path)
    for name in ("coral_py_string_list", "coral_py_logits_rows"):
        pyfn = getattr(pydll, name)
        pyfn.restype, pyfn.argtypes = SIGNATURES[name]
        setattr(lib, name, pyfn)
    _lib = lib
    return lib


def check(status: int) -> None:
    """Map a C status to the exception the reference's callers already see."""
    if status == OK:
        return
    msg = (load().coral_last_error() or b"").decode("utf-8", "replace")
    if status == EARG:
        raise ValueError(msg)
    if status == EIO:
        raise OSError(msg)
    raise CoralError(status, msg)


def ptr(t) -> int | None:
    """Device/host address of a torch tensor or numpy array (None passes NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr(device) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream
