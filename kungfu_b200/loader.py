"""Load the native runtime through its C ABI with ctypes — for foreign bindings and for programs that must not import torch.

Reference: srcs/python/kungfu/loader.py:6-34 (`_load_clib(name)` finds `<name>.so` next to the package, `_call_method(lib, name)`
calls an optional init hook, `_call_method_with`), used by `kungfu.python` to load `libkungfu_python.so` and call
`kungfu_python_init`. Here the shared object is the in-tree `_kfrt` extension: besides the pybind module it exports the C ABI of
`csrc/runtime/capi.h` (`kungfu_init`, `kungfu_all_reduce`, `kungfu_resize`, …), which `load_capi()` returns with argument types set.

    from kungfu_b200.loader import load_capi, F32, SUM
    lib = load_capi(); lib.kungfu_init()
    lib.kungfu_all_reduce(send, recv, n, F32, SUM, b"grad", None, None)
"""
from __future__ import annotations

import ctypes
import glob
import os
import platform
from ctypes import c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

F32, BF16, F16, I32, I64, U8, F64, I8, U32, U64, I16, U16 = range(12)
SUM, MIN, MAX, PROD = range(4)
CALLBACK = ctypes.CFUNCTYPE(None, c_void_p)


def _module_path() -> str:
    return os.path.dirname(os.path.abspath(__file__))


def _find(name: str) -> str:
    suffix = "so" if platform.uname()[0] != "Darwin" else "dylib"
    exact = os.path.join(_module_path(), f"{name}.{suffix}")
    if os.path.exists(exact):
        return exact
    hits = sorted(glob.glob(os.path.join(_module_path(), f"{name}.*.{suffix}")))  # _kfrt.cpython-312-x86_64-linux-gnu.so
    if not hits:
        raise FileNotFoundError(f"{name}.{suffix} not found next to {_module_path()} (build it: python -m kungfu_b200._build)")
    return hits[0]


def _load_clib(name: str) -> ctypes.CDLL:
    return ctypes.CDLL(_find(name), mode=ctypes.RTLD_GLOBAL)


def _call_method(lib, name: str, force: bool = False) -> bool:
    if hasattr(lib, name):
        getattr(lib, name)()
        return True
    if force:
        raise RuntimeError("failed to call %s" % name)
    return False


def _call_method_with(lib, name: str, *args):
    if not hasattr(lib, name):
        raise RuntimeError("failed to call %s" % name)
    return getattr(lib, name)(*args)


_SIGNATURES = {
    "kungfu_init": ([], c_int),
    "kungfu_init_single_machine": ([c_int, c_int, c_int], c_int),
    "kungfu_finalize": ([], c_int),
    "kungfu_last_error": ([], c_char_p),
    "kungfu_rank": ([], c_int),
    "kungfu_size": ([], c_int),
    "kungfu_local_rank": ([], c_int),
    "kungfu_local_size": ([], c_int),
    "kungfu_host_count": ([], c_int),
    "kungfu_uid": ([], c_uint64),
    "kungfu_detached": ([], c_int),
    "kungfu_cluster_version": ([], c_int),
    "kungfu_barrier": ([], c_int),
    "kungfu_consensus": ([c_void_p, c_size_t, c_char_p, ctypes.POINTER(c_int)], c_int),
    "kungfu_all_reduce": ([c_void_p, c_void_p, c_size_t, c_int, c_int, c_char_p, c_void_p, c_void_p], c_int),
    "kungfu_cross_all_reduce": ([c_void_p, c_void_p, c_size_t, c_int, c_int, c_char_p], c_int),
    "kungfu_reduce": ([c_void_p, c_void_p, c_size_t, c_int, c_int, c_char_p], c_int),
    "kungfu_broadcast": ([c_void_p, c_void_p, c_size_t, c_int, c_char_p, c_void_p, c_void_p], c_int),
    "kungfu_local_reduce": ([c_void_p, c_void_p, c_size_t, c_int, c_int, c_char_p], c_int),
    "kungfu_local_broadcast": ([c_void_p, c_void_p, c_size_t, c_int, c_char_p], c_int),
    "kungfu_gather": ([c_void_p, c_void_p, c_size_t, c_int, c_char_p], c_int),
    "kungfu_all_gather": ([c_void_p, c_void_p, c_size_t, c_int, c_char_p], c_int),
    "kungfu_all_reduce_int_max": ([ctypes.POINTER(c_int64)], c_int),
    "kungfu_save": ([c_char_p, c_void_p, c_size_t], c_int),
    "kungfu_save_version": ([c_char_p, c_char_p, c_void_p, c_size_t], c_int),
    "kungfu_request": ([c_int, c_char_p, c_char_p, c_void_p, c_size_t, ctypes.POINTER(c_int)], c_int),
    "kungfu_resize": ([c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)], c_int),
    "kungfu_resize_from_url": ([ctypes.POINTER(c_int), ctypes.POINTER(c_int)], c_int),
    "kungfu_propose_new_size": ([c_int], c_int),
    "kungfu_recover": ([c_double, ctypes.POINTER(c_int), ctypes.POINTER(c_int)], c_int),
    "kungfu_failure_pending": ([ctypes.POINTER(c_int)], c_int),
    "kungfu_set_tree": ([ctypes.POINTER(c_int32), c_int], c_int),
    "kungfu_get_peer_latencies": ([ctypes.POINTER(ctypes.c_float), c_int], c_int),
    "kungfu_noop": ([c_void_p, c_void_p], c_int),
}

_capi = None


def load_capi() -> ctypes.CDLL:
    """The runtime's C ABI (csrc/runtime/capi.h) as a ctypes library with argument / result types declared. Inside a process that
    already runs the Python API the handle comes from the loaded `_kfrt` extension (ONE runtime instance per process); otherwise from
    `lib/libkungfu-comm.so`, the same runtime without any dependency on Python."""
    global _capi
    if _capi is None:
        import sys

        standalone = os.path.join(_module_path(), "lib", "libkungfu-comm.so")
        if "kungfu_b200._kfrt" in sys.modules or not os.path.exists(standalone):
            lib = _load_clib("_kfrt")
        else:
            lib = ctypes.CDLL(standalone, mode=ctypes.RTLD_GLOBAL)
        for name, (args, res) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = args, res
        _capi = lib
    return _capi
