"""ctypes binding of the C ABI declared in include/sinc_b200.h (libsinc_b200.so).

No torch types cross this boundary: plain device pointers, sizes and a cudaStream_t.
The library is built in-tree by build.py; if it is missing and cannot be built the import of the
compute path fails loudly -- there is no CPU or PyTorch fallback by design.
"""
from __future__ import annotations

import ctypes
import os
import threading

from . import build as _build

DTYPE_F32 = 0
DTYPE_BF16 = 1
ENGINE_AUTO, ENGINE_FMA, ENGINE_TC, ENGINE_TC_TF32 = 0, 1, 2, 3

ABI_SYMBOLS = (
    "sinc_abi_version", "sinc_strerror", "sinc_kernel_launch_count", "sinc_set_engine",
    "sinc_get_engine", "sinc_last_engine_mask", "sinc_workspace_bytes", "sinc_synthesize",
    "sinc_forward", "sinc_backward", "sinc_tc_probe", "sinc_tc_status",
    "sinc_forward_gn", "sinc_forward_gn_partials", "sinc_gn_chunks", "sinc_gn_stats", "sinc_gn_gelu_forward",
    "sinc_gn_gelu_backward", "sinc_tc_mma_rate", "sinc_set_reserved_sms", "sinc_tc_phase_stamps",
    "sinc_mbar_wait_cost",
)


class FilterBank(ctypes.Structure):
    """struct sinc_filterbank (include/sinc_b200.h)."""
    _fields_ = [
        ("low_hz", ctypes.c_void_p),
        ("band_hz", ctypes.c_void_p),
        ("window", ctypes.c_void_p),
        ("n", ctypes.c_void_p),
        ("C", ctypes.c_int32),
        ("K", ctypes.c_int32),
        ("sample_rate", ctypes.c_double),
        ("min_low_hz", ctypes.c_double),
        ("min_band_hz", ctypes.c_double),
    ]


_lock = threading.Lock()
_lib = None


def _declare(lib):
    c = ctypes
    lib.sinc_abi_version.restype = c.c_int
    lib.sinc_abi_version.argtypes = []
    lib.sinc_strerror.restype = c.c_char_p
    lib.sinc_strerror.argtypes = [c.c_int]
    lib.sinc_kernel_launch_count.restype = c.c_uint64
    lib.sinc_kernel_launch_count.argtypes = []
    lib.sinc_set_engine.restype = c.c_int
    lib.sinc_set_engine.argtypes = [c.c_int]
    lib.sinc_set_reserved_sms.restype = c.c_int
    lib.sinc_set_reserved_sms.argtypes = [c.c_int]
    lib.sinc_get_engine.restype = c.c_int
    lib.sinc_get_engine.argtypes = []
    lib.sinc_last_engine_mask.restype = c.c_int
    lib.sinc_last_engine_mask.argtypes = []
    lib.sinc_workspace_bytes.restype = c.c_size_t
    lib.sinc_workspace_bytes.argtypes = [c.c_int64, c.c_int64, c.c_int32, c.c_int32]
    lib.sinc_synthesize.restype = c.c_int
    lib.sinc_synthesize.argtypes = [c.POINTER(FilterBank), c.c_void_p, c.c_void_p]
    lib.sinc_forward.restype = c.c_int
    lib.sinc_forward.argtypes = [c.POINTER(FilterBank), c.c_void_p, c.c_int64, c.c_void_p, c.c_int32,
                                 c.c_int64, c.c_int64, c.c_void_p, c.c_size_t, c.c_void_p]
    lib.sinc_backward.restype = c.c_int
    lib.sinc_backward.argtypes = [c.POINTER(FilterBank), c.c_void_p, c.c_int64, c.c_void_p, c.c_int32,
                                  c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
                                  c.c_int64, c.c_int64, c.c_void_p, c.c_size_t, c.c_void_p]
    lib.sinc_forward_gn_partials.restype = c.c_int64
    lib.sinc_forward_gn_partials.argtypes = [c.c_int64]
    lib.sinc_forward_gn.restype = c.c_int
    lib.sinc_forward_gn.argtypes = [c.POINTER(FilterBank), c.c_void_p, c.c_int64, c.c_void_p, c.c_int64, c.c_int64,
                                    c.c_void_p, c.c_void_p]
    lib.sinc_gn_chunks.restype = c.c_int
    lib.sinc_gn_chunks.argtypes = [c.c_int64]
    lib.sinc_gn_stats.restype = c.c_int
    lib.sinc_gn_stats.argtypes = [c.c_void_p, c.c_int64, c.c_int64, c.c_int32, c.c_void_p, c.c_void_p]
    lib.sinc_gn_gelu_forward.restype = c.c_int
    lib.sinc_gn_gelu_forward.argtypes = [c.c_void_p, c.c_void_p, c.c_int32, c.c_int32, c.c_void_p, c.c_void_p, c.c_void_p,
                                         c.c_void_p, c.c_int64, c.c_int64, c.c_int32, c.c_int32, c.c_float, c.c_void_p]
    lib.sinc_gn_gelu_backward.restype = c.c_int
    lib.sinc_gn_gelu_backward.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
                                          c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64, c.c_int64, c.c_int32, c.c_int32,
                                          c.c_void_p]
    lib.sinc_tc_mma_rate.restype = c.c_int
    lib.sinc_tc_mma_rate.argtypes = [c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_void_p, c.c_void_p]
    lib.sinc_mbar_wait_cost.restype = c.c_int
    lib.sinc_mbar_wait_cost.argtypes = [c.c_int, c.c_int, c.c_void_p, c.c_void_p]
    lib.sinc_tc_phase_stamps.restype = c.c_int
    lib.sinc_tc_phase_stamps.argtypes = [c.POINTER(c.c_longlong)]
    lib.sinc_tc_status.restype = c.c_int
    lib.sinc_tc_status.argtypes = [c.POINTER(c.c_int)]
    lib.sinc_tc_probe.restype = c.c_int
    lib.sinc_tc_probe.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_int, c.c_ulonglong, c.c_ulonglong, c.c_uint,
                                  c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_void_p,
                                  c.c_void_p]


def library_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if the .so is absent and nvcc is available) and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        if not os.path.exists(path):
            if not build_if_missing:
                raise RuntimeError(f"{path} is missing: run `python -m sincformer_metacog_speech_enhancement_b200.build`")
            _build.build_library()
        lib = ctypes.CDLL(path)
        _declare(lib)
        ver = lib.sinc_abi_version()
        if ver != 1:
            raise RuntimeError(f"libsinc_b200.so ABI version {ver}, expected 1")
        _lib = lib
        return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().sinc_strerror(code).decode()
        raise RuntimeError(f"{what} failed ({code}): {msg}")
