"""B200-native SincNet front-end: a drop-in for the reference's agents.perception.SincConv1d
(forward + backward) backed by hand-written sm_100a CUDA kernels behind a C ABI."""
from .sinc import SincConv1d, SincConvFunction, install_into_reference
from .distributed import allreduce_filter_grads, shard_rows
from .frontend import FusedGroupNormGELU, SincFrontEnd
from . import _lib


def set_engine(name: str) -> None:
    """Package default kernel family for modules constructed without ``engine=``: 'auto' (default), 'fma' (FP32
    CUDA-core kernels), 'tc' (tcgen05 tensor-core kernels; 16-bit split operands where the stated tolerance allows) or
    'tc_tf32' (tcgen05 with the 3xTF32 split in every pass).  A host-side default only: the choice travels to the C
    ABI with each call (sinc_options.engine); the library itself keeps no setting."""
    from . import sinc as _sinc
    if name not in _lib.ENGINE_CODES:
        raise ValueError(f"engine must be one of {sorted(_lib.ENGINE_CODES)}")
    _sinc._default_engine = name


def get_engine() -> str:
    from . import sinc as _sinc
    return _sinc._default_engine


def kernel_launch_count() -> int:
    return int(_lib.load().sinc_kernel_launch_count())


__all__ = ["SincConv1d", "SincConvFunction", "install_into_reference", "allreduce_filter_grads",
           "shard_rows", "set_engine", "get_engine", "kernel_launch_count", "FusedGroupNormGELU", "SincFrontEnd"]
