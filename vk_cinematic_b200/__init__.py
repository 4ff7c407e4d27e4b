"""vk_cinematic_b200 -- B200 (sm_100a) implementation of vk_cinematic's CPU "SIMD" path tracer
(the sp_ path) behind the reference's own sp_ interface.

Contents: csrc/ (CUDA kernels, host builder, the C ABI of libspb200.so), sp.py (ctypes mirror
of the sp_ interface), workloads.py (synthetic inputs for the BASELINE configurations),
strips.py (tile-strip partition across GPUs).  Importing the package loads libspb200.so and
raises if it, or any symbol include/sp_b200.h declares, is missing -- there is no CPU path.
"""
from . import sp  # noqa: F401  (loads and checks the shared library)
from . import workloads  # noqa: F401
from . import strips  # noqa: F401

__all__ = ["sp", "workloads", "strips"]
