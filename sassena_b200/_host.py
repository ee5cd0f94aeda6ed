"""ctypes signatures of the host-layer C API (include/sassena_host.h)."""
import ctypes as C

c_size_p = C.POINTER(C.c_size_t)
c_double_p = C.POINTER(C.c_double)
c_long_p = C.POINTER(C.c_long)

RANK_FN = C.CFUNCTYPE(C.c_size_t, C.c_void_p)
SIZE_FN = C.CFUNCTYPE(C.c_size_t, C.c_void_p)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)
BARRIER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)
SPLIT_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int)
RELEASE_FN = C.CFUNCTYPE(None, C.c_void_p)
NCCL_COMM_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p)


class CommVtbl(C.Structure):
    _fields_ = [("user", C.c_void_p), ("rank", RANK_FN), ("size", SIZE_FN), ("allreduce_sum", ALLREDUCE_FN),
                ("barrier", BARRIER_FN), ("split", SPLIT_FN), ("release", RELEASE_FN), ("nccl_comm", NCCL_COMM_FN)]


# backend table: same order as sass_backend_vtbl
BE_INIT = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_void_p))
BE_DESTROY = C.CFUNCTYPE(None, C.c_void_p)
BE_LAST_ERROR = C.CFUNCTYPE(C.c_char_p, C.c_void_p)
BE_SYNC = C.CFUNCTYPE(C.c_int, C.c_void_p)
BE_STAGE_FRAMES = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int)
BE_TO_SPH = C.CFUNCTYPE(C.c_int, C.c_void_p)
BE_STAGE_ATOMS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t)
BE_STAGE_ATOMS_FF = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t)
BE_SET_FACTORS = C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, C.c_size_t)
BE_PARTIAL_LEN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, c_size_p)
BE_COMPUTE_VEC = C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, C.c_size_t, C.c_int, C.c_void_p)
BE_COMPUTE_MP = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_double, c_long_p, C.c_size_t, C.c_int, C.c_void_p)
BE_FINALIZE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, c_double_p, c_double_p,
                          c_double_p)
BE_SET_FACTORS_BATCH = C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, C.c_size_t, C.c_size_t)
BE_MP_AMPL = C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, C.c_size_t, c_long_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p)
BE_MP_DSP = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p)
BE_SET_WINDOW = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_size_t)
BE_AV_AMPL = C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, C.c_size_t, C.c_void_p)
BE_AV_DSP = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p)
BE_AV_SCAN = C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, C.c_size_t, c_double_p, C.c_size_t, C.c_int, C.c_void_p)
BE_AV_SCAN_AMPL = C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, C.c_size_t, c_double_p, C.c_size_t, C.c_void_p)
BE_STAGE_WAVE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t)
BE_ACCUMULATE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)
BE_TO_CYL = C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p)
BE_CYL_AMPL = C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, c_double_p, c_long_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p)
BE_PREFETCH = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t)
BE_SWAP = C.CFUNCTYPE(C.c_int, C.c_void_p)
BE_ALLOC = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_void_p), C.c_size_t)
BE_FREE = C.CFUNCTYPE(C.c_int, C.c_void_p)


class BackendVtbl(C.Structure):
    _fields_ = [("init", BE_INIT), ("destroy", BE_DESTROY), ("last_error", BE_LAST_ERROR), ("synchronize", BE_SYNC),
                ("stage_frames", BE_STAGE_FRAMES), ("frames_to_spherical", BE_TO_SPH), ("stage_atoms", BE_STAGE_ATOMS),
                ("stage_atoms_from_frames", BE_STAGE_ATOMS_FF), ("set_factors", BE_SET_FACTORS),
                ("partial_len", BE_PARTIAL_LEN), ("compute_all_vectors_partial", BE_COMPUTE_VEC),
                ("compute_self_vectors_partial", BE_COMPUTE_VEC), ("compute_mpsphere_partial", BE_COMPUTE_MP),
                ("finalize", BE_FINALIZE), ("device_alloc", BE_ALLOC), ("device_free", BE_FREE),
                ("set_factors_batch", BE_SET_FACTORS_BATCH), ("mpsphere_amplitudes", BE_MP_AMPL),
                ("mpsphere_dsp_partial", BE_MP_DSP), ("set_frame_window", BE_SET_WINDOW),
                ("all_vectors_amplitudes", BE_AV_AMPL), ("all_vectors_dsp_partial", BE_AV_DSP),
                ("compute_all_vectors_scan_partial", BE_AV_SCAN), ("all_vectors_scan_amplitudes", BE_AV_SCAN_AMPL),
                ("stage_atoms_wave", BE_STAGE_WAVE), ("accumulate", BE_ACCUMULATE),
                ("frames_to_cylindrical", BE_TO_CYL), ("mpcylinder_amplitudes", BE_CYL_AMPL),
                ("stage_atoms_prefetch", BE_PREFETCH), ("stage_atoms_swap", BE_SWAP),
                ("host_alloc", BE_ALLOC), ("host_free", BE_FREE),
                ("comm_adopt", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int)),
                ("compute_all_vectors_scan_sharded",
                 C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, C.c_size_t, c_double_p, C.c_size_t, C.c_int, C.c_void_p)),
                ("compute_all_vectors_sharded", C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, C.c_size_t, C.c_int, C.c_void_p))]


FACTORS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double, c_double_p, C.c_size_t)
WRITE_FN = C.CFUNCTYPE(None, C.c_void_p, c_double_p, c_double_p, C.c_size_t, c_double_p, c_double_p)

SIGNATURES = {
    "sass_comm_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "sass_comm_nccl_bootstrap_file": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_char_p]),
    "sass_comm_nccl_create": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sass_last_error": (C.c_char_p, []),
    "sass_params_new": (C.c_void_p, []),
    "sass_params_free": (None, [C.c_void_p]),
    "sass_params_set": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "sass_params_set_vectors": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t]),
    "sass_params_set_moments": (C.c_int, [C.c_void_p, c_long_p, C.c_size_t]),
    "sass_params_create": (C.c_int, [C.c_void_p]),
    "sass_params_num_vectors": (C.c_size_t, [C.c_void_p]),
    "sass_params_get_vectors": (C.c_int, [C.c_void_p, c_double_p]),
    "sass_params_num_moments": (C.c_size_t, [C.c_void_p]),
    "sass_params_get_moments": (C.c_int, [C.c_void_p, c_long_p]),
    "sass_scatter_run": (C.c_int, [C.c_void_p, C.POINTER(CommVtbl), C.POINTER(BackendVtbl), C.c_void_p, C.c_size_t,
                                   C.c_size_t, C.c_void_p, c_double_p, FACTORS_FN, C.c_void_p, c_double_p, C.c_size_t,
                                   WRITE_FN, C.c_void_p, C.POINTER(C.c_int), C.c_char_p, C.c_size_t]),
    "sass_motion_transforms": (C.c_int, [C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_ulong, C.c_long,
                                         C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_double)]),
    "sass_div_assignment": (C.c_int, [C.c_size_t, C.c_size_t, C.c_size_t, c_size_p, c_size_p, c_size_p]),
    "sass_mod_assignment": (C.c_int, [C.c_size_t, C.c_size_t, C.c_size_t, c_size_p, c_size_p, c_size_p]),
    "sass_decomposition_plan": (C.c_int, [C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_double,
                                          C.c_int, C.c_size_t, c_size_p, c_size_p, c_size_p]),
    "sass_create_from_scans": (C.c_size_t, [c_double_p, C.c_size_t, c_double_p, C.c_size_t]),
    "sass_dcd_open": (C.c_int, [C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "sass_dcd_info": (C.c_int, [C.c_void_p, c_size_p, c_size_p, C.POINTER(C.c_int)]),
    "sass_dcd_read": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "sass_dcd_close": (None, [C.c_void_p]),
    "sass_dcd_write": (C.c_int, [C.c_char_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "sass_xdr_open": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "sass_xdr_info": (C.c_int, [C.c_void_p, c_size_p, c_size_p]),
    "sass_xdr_read": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "sass_xdr_close": (None, [C.c_void_p]),
    "sass_h5_read": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p]),
    "sass_h5_write_signal": (C.c_int, [C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p,
                                       C.c_size_t, c_double_p, c_double_p, c_double_p, c_double_p, c_size_p]),
    "sass_init_subvectors": (C.c_size_t, [C.c_void_p, c_double_p, c_double_p, C.c_size_t]),
    "sass_job_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "sass_job_load_overwrite": (C.c_int, [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_size_t,
                                          C.POINTER(C.c_void_p)]),
    "sass_job_free": (None, [C.c_void_p]),
    "sass_job_signal_file": (C.c_char_p, [C.c_void_p]),
    "sass_job_option": (C.c_char_p, [C.c_void_p, C.c_char_p]),
    "sass_job_info": (C.c_int, [C.c_void_p, c_size_p, c_size_p, c_size_p, c_size_p]),
    "sass_job_qvectors": (C.c_int, [C.c_void_p, c_double_p]),
    "sass_job_factors": (C.c_int, [C.c_void_p, C.c_double, c_double_p]),
    "sass_job_frames": (C.c_int, [C.c_void_p, C.POINTER(C.POINTER(C.c_float))]),
    "sass_job_selection": (C.c_int, [C.c_void_p, C.c_char_p, c_size_p, C.c_size_t, c_size_p]),
    "sass_job_params": (C.c_void_p, [C.c_void_p]),
    "sass_job_run": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(CommVtbl), C.POINTER(BackendVtbl), C.c_void_p, c_size_p,
                               C.c_char_p, C.c_size_t]),
    "sass_job_stage": (C.c_int, [C.c_void_p, C.POINTER(CommVtbl), C.POINTER(BackendVtbl), C.c_void_p, c_size_p, C.c_char_p,
                                 C.c_size_t]),
}


def bind(lib):
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
