"""sassena_b200 — B200-native (sm_100a) implementation of the Sassena scattering hot path.

The product is the C-ABI shared library ``libsassena_b200.so`` (include/sassena_b200.h): hand-written CUDA
kernels for the amplitude, orientational-average/multipole and FFT-autocorrelation steps plus the C++ host
layer mirroring the reference's scatter_devices / stager / decomposition interface.  This Python package is
a thin ctypes binding used by tests and bench.py; it contains no compute and no CPU fallback.
"""
from ._lib import load_library, LibraryMissing  # noqa: F401
from .api import (  # noqa: F401
    DSP_AUTOCORRELATE, DSP_PLAIN, DSP_SQUARE, METHOD_DIRECT, METHOD_FFTW, REPR_CARTESIAN, REPR_SPHERICAL,
    ScatterContext, SgpuError, comm_unique_id,
)
