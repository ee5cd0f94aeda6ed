"""ctypes loader for libsassena_b200.so.  Fails loudly when the CUDA extension is missing."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsassena_b200.so")


class LibraryMissing(RuntimeError):
    pass


_lib = None

c_size_p = C.POINTER(C.c_size_t)
c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_long_p = C.POINTER(C.c_long)

# name -> (restype, argtypes); mirrors include/sassena_b200.h one to one
SIGNATURES = {
    "sgpu_init": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "sgpu_destroy": (None, [C.c_void_p]),
    "sgpu_last_error": (C.c_char_p, [C.c_void_p]),
    "sgpu_version": (C.c_char_p, []),
    "sgpu_launch_count": (C.c_uint64, [C.c_void_p]),
    "sgpu_synchronize": (C.c_int, [C.c_void_p]),
    "sgpu_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "sgpu_host_free": (C.c_int, [C.c_void_p]),
    "sgpu_stage_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]),
    "sgpu_stage_frames_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]),
    "sgpu_frames_to_spherical": (C.c_int, [C.c_void_p]),
    "sgpu_frames_to_cylindrical": (C.c_int, [C.c_void_p, c_double_p]),
    "sgpu_compute_mpcylinder": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_long_p, C.c_size_t, C.c_int, C.c_int, c_double_p,
                                          c_double_p, c_double_p]),
    "sgpu_compute_mpcylinder_partial": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_long_p, C.c_size_t, C.c_int, C.c_void_p]),
    "sgpu_mpcylinder_amplitudes": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_long_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                             C.c_void_p]),
    "sgpu_stage_atoms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "sgpu_stage_atoms_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "sgpu_stage_atoms_from_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]),
    "sgpu_comm_get_unique_id": (C.c_int, [C.c_char_p]),
    "sgpu_comm_init": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "sgpu_comm_adopt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "sgpu_comm_destroy": (C.c_int, [C.c_void_p]),
    "sgpu_comm_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sgpu_comm_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "sgpu_compute_all_vectors_scan_sharded": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t, c_double_p, C.c_size_t, C.c_int, C.c_void_p]),
    "sgpu_compute_all_vectors_sharded": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t, C.c_int, C.c_void_p]),
    "sgpu_stage_atoms_prefetch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "sgpu_stage_atoms_swap": (C.c_int, [C.c_void_p]),
    "sgpu_staged_shape": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "sgpu_device_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "sgpu_stage_atoms_wave": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]),
    "sgpu_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "sgpu_set_factors": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t]),
    "sgpu_compute_all_vectors": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]),
    "sgpu_compute_self_vectors": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]),
    "sgpu_compute_mpsphere": (C.c_int, [C.c_void_p, C.c_double, c_long_p, C.c_size_t, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]),
    "sgpu_set_factors_batch": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t, C.c_size_t]),
    "sgpu_compute_mpsphere_batch": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t, c_long_p, C.c_size_t, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]),
    "sgpu_compute_mpsphere_batch_partial": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t, c_long_p, C.c_size_t, C.c_int, C.c_void_p]),
    "sgpu_mpsphere_amplitudes": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t, c_long_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]),
    "sgpu_set_frame_window": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "sgpu_all_vectors_amplitudes": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_size_t, C.c_void_p]),
    "sgpu_all_vectors_dsp_partial": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]),
    "sgpu_compute_all_vectors_scan": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_double),
                                               C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double),
                                               C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "sgpu_compute_all_vectors_scan_partial": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_size_t,
                                                       C.POINTER(C.c_double), C.c_size_t, C.c_int, C.c_void_p]),
    "sgpu_all_vectors_scan_amplitudes": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_double),
                                                  C.c_size_t, C.c_void_p]),
    "sgpu_last_scan_plan": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sgpu_mpsphere_dsp_partial": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]),
    "sgpu_partial_len": (C.c_int, [C.c_void_p, C.c_int, c_size_p]),
    "sgpu_compute_all_vectors_partial": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t, C.c_int, C.c_void_p]),
    "sgpu_compute_self_vectors_partial": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t, C.c_int, C.c_void_p]),
    "sgpu_compute_mpsphere_partial": (C.c_int, [C.c_void_p, C.c_double, c_long_p, C.c_size_t, C.c_int, C.c_void_p]),
    "sgpu_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, c_double_p, c_double_p, c_double_p]),
    "sgpu_stream": (C.c_void_p, [C.c_void_p]),
    "sgpu_get_amplitudes": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t, C.c_size_t]),
    "sgpu_last_amplitude_ms": (C.c_int, [C.c_void_p, c_float_p]),
    "sgpu_last_dsp_ms": (C.c_int, [C.c_void_p, c_float_p]),
    "sgpu_timer_start": (C.c_int, [C.c_void_p]),
    "sgpu_timer_stop": (C.c_int, [C.c_void_p, c_float_p]),
    "sgpu_measure_fp64_peak": (C.c_int, [C.c_void_p, c_double_p]),
    "sgpu_synth_trajectory": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_uint64, C.c_int]),
    "sgpu_device_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "sgpu_device_free": (C.c_int, [C.c_void_p]),
    "sgpu_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "sgpu_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
}


def load_library(path: str | None = None):
    """Load the shared library and bind every symbol of the header.  Raises LibraryMissing if the .so has not
    been built (python -m sassena_b200.build) — there is deliberately no pure-Python or CPU fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("SASSENA_B200_LIB") or LIB_PATH  # SASSENA_B200_LIB: an alternative build of the library
    if not os.path.exists(p):
        raise LibraryMissing(
            f"{p} not found: build the CUDA extension first (python -m sassena_b200.build or __graft_entry__.build()); "
            "sassena_b200 has no CPU fallback")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    try:
        from . import _host  # C++ host-layer symbols (sass_*), optional until built
        _host.bind(lib)
    except ImportError:
        pass
    if path is None:
        _lib = lib
    return lib
