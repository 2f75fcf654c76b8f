"""ctypes binding of libcape_msda.so (C ABI in include/cape_msda.h).

There is deliberately no fallback: if the shared object is missing or fails to load, every op raises
``CapeLibraryError`` telling the user to run ``python __graft_entry__.py build`` — nothing silently routes to
PyTorch eager, the CPU or the test oracle.
"""
from __future__ import annotations

import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# CAPE_MSDA_LIB points the loader at another build of the same ABI (used by tools/ for profiling variants).
LIB_PATH = os.environ.get("CAPE_MSDA_LIB") or os.path.join(HERE, "libcape_msda.so")
ABI_VERSION = 1

DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2


class CapeLibraryError(RuntimeError):
    pass


class Dims(ctypes.Structure):
    """``cape_msda_dims`` (include/cape_msda.h)."""
    _fields_ = [("N", ctypes.c_int32), ("S", ctypes.c_int32), ("M", ctypes.c_int32), ("D", ctypes.c_int32),
                ("Lq", ctypes.c_int32), ("L", ctypes.c_int32), ("P", ctypes.c_int32)]


class Tokenizer(ctypes.Structure):
    """``cape_tokenizer`` (include/cape_msda.h)."""
    _fields_ = [("num_bins", ctypes.c_int32), ("min_len", ctypes.c_int32), ("bos", ctypes.c_int64),
                ("eos", ctypes.c_int64), ("sep", ctypes.c_int64), ("pad", ctypes.c_int64), ("cls", ctypes.c_int64),
                ("type_coord", ctypes.c_int32), ("type_sep", ctypes.c_int32), ("type_eos", ctypes.c_int32),
                ("type_cls", ctypes.c_int32)]


class TokenState(ctypes.Structure):
    """``cape_token_state`` (include/cape_msda.h): device pointers of the per-sample generation buffers."""
    _fields_ = [(name, ctypes.c_void_p) for name in (
        "unfinished", "finish_step", "seq11", "seq12", "seq21", "seq22", "delta_x1", "delta_x2", "delta_y1", "delta_y2",
        "pred_logits", "pred_coords", "gen_kind", "gen_xy")] + [("max_len", ctypes.c_int64)]


_lock = threading.Lock()
_lib = None
_load_error = None

_vp, _i = ctypes.c_void_p, ctypes.c_int
_SIGNATURES = {
    "cape_abi_version": (ctypes.c_int, []),
    "cape_last_error": (ctypes.c_char_p, []),
    "cape_launch_count": (ctypes.c_uint64, []),
    "cape_set_tuning": (_i, [ctypes.c_char_p, _i]),
    "cape_get_tuning": (_i, [ctypes.c_char_p]),
    "cape_debug_counters": (_i, [ctypes.POINTER(ctypes.c_longlong), _i]),
    "cape_msda_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(Dims), _i, _i, _vp]),
    "cape_msda_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(Dims), _i, _i, _i, _vp]),
    "cape_msda_decode": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(Dims), _i, _vp]),
    "cape_msda_fused_supported": (_i, [ctypes.POINTER(Dims)]),
    "cape_msda_fused_backward": (_i, [_vp] * 10 + [ctypes.POINTER(Dims), _i, _i, _vp]),
    "cape_msda_query_pool_forward": (_i, [_vp] * 6 + [ctypes.POINTER(Dims), _vp]),
    "cape_msda_query_pool_backward": (_i, [_vp] * 9 + [ctypes.POINTER(Dims), _i, _vp]),
    "cape_points_sample_forward": (_i, [_vp] * 3 + [_i] * 7 + [_vp]),
    "cape_points_sample_backward": (_i, [_vp] * 5 + [_i] * 8 + [_vp]),
    "cape_zero_masked_rows": (_i, [_vp, _vp, ctypes.c_int64, _i, _vp]),
    "cape_seq_embed_forward": (_i, [_vp] * 10 + [ctypes.c_int64, _i, _i, _vp]),
    "cape_seq_embed_backward": (_i, [_vp] * 10 + [ctypes.c_int64, _i, _i, ctypes.c_int64, _i, _vp]),
    "cape_token_step": (_i, [_vp, _vp, _vp, ctypes.POINTER(TokenState), ctypes.POINTER(Tokenizer), _i, _i, _vp]),
    "cape_decode_attention": (_i, [_vp, _i, _vp, _vp, _i] + [_vp] * 5 + [_i] * 4 + [_vp]),
    "cape_skinny_linear": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, ctypes.c_float, _vp, _vp, _i, _i, _i, _i, _i,
                                _vp]),
    "cape_msda_output_proj": (_i, [_vp] * 6 + [ctypes.POINTER(Dims)] + [_vp, _vp, _vp, _i, _vp, _vp, ctypes.c_float, _vp, _i, _i, _vp]),
    "cape_skinny_linear_split": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "cape_coord_head_refine": (_i, [_vp, _i] + [_vp] * 8 + [_i] * 4 + [_vp]),
    "cape_tiny_linear": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "cape_tf32_split_lo": (_i, [_vp, _vp, ctypes.c_int64, _vp]),
    "cape_linear_tf32x3": (_i, [_vp] * 5 + [_i] * 4 + [_vp]),
    "cape_linear_tf32x3_wgrad": (_i, [_vp] * 4 + [_i] * 3 + [_vp]),
    "cape_msda_host_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(Dims), _i]),
    "cape_msda_forward_backward_host": (_i, [_vp] * 10 + [ctypes.POINTER(Dims), _vp, ctypes.c_size_t, _vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load():
    """Load (once) and return the ctypes handle; raises CapeLibraryError when the library is unavailable."""
    global _lib, _load_error
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            _load_error = (f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
                           f"(nvcc, sm_100a). There is no CPU or eager fallback.")
            raise CapeLibraryError(_load_error)
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as e:                                    # pragma: no cover - depends on the box
            _load_error = f"cannot load {LIB_PATH}: {e}"
            raise CapeLibraryError(_load_error) from e
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        got = lib.cape_abi_version()
        if got != ABI_VERSION:
            raise CapeLibraryError(f"{LIB_PATH} has ABI version {got}, this package expects {ABI_VERSION}: rebuild")
        _lib = lib
    return _lib


def available() -> bool:
    try:
        load()
        return True
    except CapeLibraryError:
        return False


def check(rc: int, what: str) -> None:
    """Turn a non-zero ABI return code into a RuntimeError carrying cape_last_error()."""
    if rc != 0:
        msg = load().cape_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().cape_launch_count())


def set_tuning(name: str, value: int) -> None:
    """Override a launch-geometry / kernel-selection knob at run time (``CAPE_<NAME>`` environment variables are read
    once per process by the library); ``value <= 0`` restores the default."""
    check(load().cape_set_tuning(name.encode(), int(value)), f"cape_set_tuning({name})")
