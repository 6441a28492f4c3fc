"""ctypes binding of include/pantas_aug.h.  Fails loudly: there is no fallback."""
from __future__ import annotations

import ctypes
import os

from .errors import NativeLibraryError

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libpantas_aug.so")

c_ctx = ctypes.c_void_p
u64 = ctypes.c_uint64
i64 = ctypes.c_int64
vp = ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol include/pantas_aug.h declares
SIGNATURES = {
    "pt_abi_version": (ctypes.c_int, []),
    "pt_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "pt_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_ctx)]),
    "pt_destroy": (None, [c_ctx]),
    "pt_last_error": (ctypes.c_char_p, [c_ctx]),
    "pt_set_stream": (ctypes.c_int, [c_ctx, vp]),
    "pt_set_graph": (ctypes.c_int, [c_ctx, vp, u64, ctypes.c_uint32, vp, u64, u64, u64]),
    "pt_reset_counts": (ctypes.c_int, [c_ctx]),
    "pt_process_chunk": (ctypes.c_int, [c_ctx, vp, u64, u64, i64]),
    "pt_process_host": (i64, [c_ctx, vp, u64, u64, i64]),
    "pt_wait_copy": (ctypes.c_int, [c_ctx, i64]),
    "pt_set_stage_bytes": (ctypes.c_int, [c_ctx, u64]),
    "pt_stage_bytes": (u64, [c_ctx]),
    "pt_sync": (ctypes.c_int, [c_ctx]),
    "pt_error": (ctypes.c_int, [c_ctx, ctypes.POINTER(u64), ctypes.POINTER(ctypes.c_int)]),
    "pt_finalize": (ctypes.c_int, [c_ctx, ctypes.POINTER(u64), ctypes.POINTER(u64)]),
    "pt_export_dense": (ctypes.c_int, [c_ctx, vp, u64, vp, u64]),
    "pt_export_side": (ctypes.c_int, [c_ctx, vp, u64, vp, u64]),
    "pt_timer_start": (ctypes.c_int, [c_ctx]),
    "pt_timer_stop": (ctypes.c_int, [c_ctx, ctypes.POINTER(ctypes.c_float)]),
    "pt_profile_enable": (ctypes.c_int, [c_ctx, ctypes.c_int]),
    "pt_kernel_time": (ctypes.c_int, [c_ctx, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(u64)]),
    "pt_kernel_time_split": (ctypes.c_int, [c_ctx, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(u64)]),
    "pt_debug_counters": (ctypes.c_int, [c_ctx, ctypes.POINTER(u64), ctypes.c_int]),
    "pt_stats": (ctypes.c_int, [c_ctx, ctypes.POINTER(u64), ctypes.POINTER(u64), ctypes.POINTER(u64)]),
    "pt_gfa_parse": (ctypes.c_int, [c_ctx, vp, vp, vp, u64, vp, vp, vp, vp, vp, vp]),
    "pt_gfa_measure": (ctypes.c_int, [c_ctx, vp, vp, vp]),
    "pt_gfa_format": (ctypes.c_int, [c_ctx, vp, vp, vp, vp]),
}


class GfaWriterArgs(ctypes.Structure):
    """struct pt_gfa_writer of include/pantas_aug.h"""
    _fields_ = [("gfa", vp), ("start", vp), ("a_rel", vp), ("slen", vp), ("kind", vp), ("v1", vp), ("link_edge", vp),
                ("node_len", vp), ("sums", vp), ("sp_slot", vp), ("sp_off", vp), ("sp_text", vp),
                ("n_nodes", u64), ("n_lines", u64), ("min_id", ctypes.c_uint32)]

_lib = None


def load_library(path: str | None = None):
    """dlopen libpantas_aug.so and type every entry point."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise NativeLibraryError(
            f"{p} not found: build it with `python -m pantas_b200.build` (nvcc, sm_100a). "
            "pantas_b200 has no CPU fallback.")
    try:
        lib = ctypes.CDLL(p)
    except OSError as e:
        raise NativeLibraryError(f"cannot load {p}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{p} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.pt_abi_version() != 1:
        raise NativeLibraryError("libpantas_aug.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib
