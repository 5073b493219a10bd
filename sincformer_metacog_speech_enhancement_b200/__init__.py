"""B200-native SincNet front-end: a drop-in for the reference's agents.perception.SincConv1d
(forward + backward) backed by hand-written sm_100a CUDA kernels behind a C ABI."""
from .sinc import SincConv1d, SincConvFunction, install_into_reference
from .distributed import allreduce_filter_grads, shard_rows
from .frontend import FusedGroupNormGELU, SincFrontEnd
from . import _lib


def set_engine(name: str) -> None:
    """'auto' (default), 'fma' (FP32 CUDA-core kernels), 'tc' (tcgen05 tensor-core kernels; 16-bit split operands
    where the stated tolerance allows) or 'tc_tf32' (tcgen05 with the 3xTF32 split in every pass)."""
    code = {"auto": _lib.ENGINE_AUTO, "fma": _lib.ENGINE_FMA, "tc": _lib.ENGINE_TC, "tc_tf32": _lib.ENGINE_TC_TF32}[name]
    _lib.check(_lib.load().sinc_set_engine(code), "sinc_set_engine")


def kernel_launch_count() -> int:
    return int(_lib.load().sinc_kernel_launch_count())


__all__ = ["SincConv1d", "SincConvFunction", "install_into_reference", "allreduce_filter_grads",
           "shard_rows", "set_engine", "kernel_launch_count", "FusedGroupNormGELU", "SincFrontEnd"]
