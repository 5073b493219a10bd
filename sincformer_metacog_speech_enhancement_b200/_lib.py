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

ABI_VERSION = 2

# every symbol include/sinc_b200.h declares
ABI_SYMBOLS = (
    "sinc_abi_version", "sinc_strerror", "sinc_kernel_launch_count", "sinc_workspace_bytes", "sinc_synthesize",
    "sinc_forward", "sinc_backward",
    "sinc_forward_gn", "sinc_forward_gn_partials", "sinc_gn_chunks", "sinc_gn_stats", "sinc_gn_gelu_forward",
    "sinc_gn_gelu_backward",
)
# every symbol include/sinc_b200_diag.h declares (libsinc_b200_diag.so: tests and tools only)
DIAG_SYMBOLS = ("sinc_tc_probe", "sinc_tc_mma_rate", "sinc_mbar_wait_cost", "sinc_tc_ld16x256_probe", "sinc_feed_probe")


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


class Options(ctypes.Structure):
    """struct sinc_options (include/sinc_b200.h): per-call settings; the library keeps no state between calls."""
    _fields_ = [
        ("engine", ctypes.c_int32),
        ("reserved_sms", ctypes.c_int32),
        ("engines_used", ctypes.POINTER(ctypes.c_int32)),
        ("phase_stamps", ctypes.c_void_p),
    ]


ENGINE_CODES = {"auto": ENGINE_AUTO, "fma": ENGINE_FMA, "tc": ENGINE_TC, "tc_tf32": ENGINE_TC_TF32}


def decode_engines(mask: int) -> dict:
    """sinc_options.engines_used -> {'fwd' | 'bwd_input' | 'bwd_filter': 'fma' | 'tc_tf32' | 'tc16'} for the passes run."""
    out = {}
    for name, tc_bit, h16_bit in (("fwd", 1, 8), ("bwd_input", 2, 16), ("bwd_filter", 4, 32)):
        out[name] = "fma" if not (mask & tc_bit) else ("tc16" if (mask & h16_bit) else "tc_tf32")
    return out


_lock = threading.Lock()
_lib = None
_diag = None


def _declare(lib):
    c = ctypes
    lib.sinc_abi_version.restype = c.c_int
    lib.sinc_abi_version.argtypes = []
    lib.sinc_strerror.restype = c.c_char_p
    lib.sinc_strerror.argtypes = [c.c_int]
    lib.sinc_kernel_launch_count.restype = c.c_uint64
    lib.sinc_kernel_launch_count.argtypes = []
    lib.sinc_workspace_bytes.restype = c.c_size_t
    lib.sinc_workspace_bytes.argtypes = [c.c_int64, c.c_int64, c.c_int32, c.c_int32]
    lib.sinc_synthesize.restype = c.c_int
    lib.sinc_synthesize.argtypes = [c.POINTER(FilterBank), c.c_void_p, c.c_void_p]
    lib.sinc_forward.restype = c.c_int
    lib.sinc_forward.argtypes = [c.POINTER(FilterBank), c.c_void_p, c.c_int64, c.c_void_p, c.c_int32,
                                 c.c_int64, c.c_int64, c.c_void_p, c.c_size_t, c.c_void_p, c.POINTER(Options)]
    lib.sinc_backward.restype = c.c_int
    lib.sinc_backward.argtypes = [c.POINTER(FilterBank), c.c_void_p, c.c_int64, c.c_void_p, c.c_int32,
                                  c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
                                  c.c_int64, c.c_int64, c.c_void_p, c.c_size_t, c.c_void_p, c.POINTER(Options)]
    lib.sinc_forward_gn_partials.restype = c.c_int64
    lib.sinc_forward_gn_partials.argtypes = [c.c_int64]
    lib.sinc_forward_gn.restype = c.c_int
    lib.sinc_forward_gn.argtypes = [c.POINTER(FilterBank), c.c_void_p, c.c_int64, c.c_void_p, c.c_int64, c.c_int64,
                                    c.c_void_p, c.c_void_p, c.POINTER(Options)]
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


def _declare_diag(lib):
    c = ctypes
    lib.sinc_tc_mma_rate.restype = c.c_int
    lib.sinc_tc_mma_rate.argtypes = [c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_void_p, c.c_void_p]
    lib.sinc_mbar_wait_cost.restype = c.c_int
    lib.sinc_mbar_wait_cost.argtypes = [c.c_int, c.c_int, c.c_void_p, c.c_void_p]
    lib.sinc_tc_probe.restype = c.c_int
    lib.sinc_tc_probe.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_int, c.c_ulonglong, c.c_ulonglong, c.c_uint,
                                  c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_void_p,
                                  c.c_void_p]
    lib.sinc_tc_ld16x256_probe.restype = c.c_int
    lib.sinc_tc_ld16x256_probe.argtypes = [c.c_void_p, c.c_void_p]
    lib.sinc_feed_probe.restype = c.c_int
    lib.sinc_feed_probe.argtypes = [c.c_void_p, c.c_int64, c.c_int32, c.c_int64, c.c_int, c.c_int, c.c_void_p, c.c_void_p]


def library_path() -> str:
    return _build.LIB_PATH


def _ensure_built(path, build_fn, fresh_fn, build_if_missing):
    """The library must exist AND belong to the sources in csrc/: a stale .so (edited kernels, another checkout) is
    rebuilt when nvcc is here and refused when it is not -- the ABI version alone cannot see a kernel change."""
    if os.path.exists(path) and fresh_fn():
        return
    if _build.find_nvcc() is not None and build_if_missing:
        build_fn()
        return
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m sincformer_metacog_speech_enhancement_b200.build`")
    raise RuntimeError(f"{path} is stale (csrc/ or include/ changed since it was built) and nvcc is not available: "
                       "rebuild with `python -m sincformer_metacog_speech_enhancement_b200.build`")


def load(build_if_missing: bool = True):
    """Load (building first if the .so is absent or stale and nvcc is available) and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        _ensure_built(path, _build.build_library, _build.is_fresh, build_if_missing)
        lib = ctypes.CDLL(path)
        _declare(lib)
        ver = lib.sinc_abi_version()
        if ver != ABI_VERSION:
            raise RuntimeError(f"libsinc_b200.so ABI version {ver}, expected {ABI_VERSION}")
        _lib = lib
        return lib


def load_diag(build_if_missing: bool = True):
    """The diagnostics library (hardware probes, micro-benchmarks; include/sinc_b200_diag.h).  Tests and tools only:
    nothing in the product path calls this."""
    global _diag
    if _diag is not None:
        return _diag
    with _lock:
        if _diag is not None:
            return _diag
        path = _build.DIAG_LIB_PATH
        _ensure_built(path, _build.build_diag_library, lambda: _build.is_fresh(diag=True), build_if_missing)
        lib = ctypes.CDLL(path)
        _declare_diag(lib)
        _diag = lib
        return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().sinc_strerror(code).decode()
        hint = ""
        if code > 0:
            hint = (" -- a CUDA error; if it is sticky ('unspecified launch failure' / 'trap'), a tcgen05 pipeline barrier "
                    "of an earlier launch timed out and the kernel trapped rather than return invalid results")
        raise RuntimeError(f"{what} failed ({code}): {msg}{hint}")
