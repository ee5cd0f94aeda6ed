"""Thin Python mirror of the C-ABI (include/sassena_b200.h).  No compute happens here."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import load_library

DSP_AUTOCORRELATE, DSP_SQUARE, DSP_PLAIN = 0, 1, 2
METHOD_FFTW, METHOD_DIRECT = 0, 1
REPR_CARTESIAN, REPR_SPHERICAL = 0, 1

_DSP = {"autocorrelate": 0, "square": 1, "plain": 2}
_METHOD = {"fftw": 0, "direct": 1}


class SgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sgpu error {code}: {msg}")
        self.code = code
        self.msg = msg


def _dsp(v):
    return _DSP[v] if isinstance(v, str) else int(v)


def _method(v):
    return _METHOD[v] if isinstance(v, str) else int(v)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class PinnedArray:
    """numpy view over cudaHostAlloc'ed memory (the stager's pinned buffers)."""

    def __init__(self, lib, shape, dtype):
        self._lib = lib
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        rc = lib.sgpu_host_alloc(C.byref(p), nbytes)
        if rc:
            raise SgpuError(rc, "sgpu_host_alloc failed")
        self.ptr = p.value
        buf = (C.c_char * nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.sgpu_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ScatterContext:
    """One GPU worth of the hot path: stage -> set_factors -> compute_* (one call per |q|)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.sgpu_init(int(device), C.byref(h))
        if rc:
            raise SgpuError(rc, self.lib.sgpu_last_error(None).decode())
        self.h = h
        self.device = device
        self.NF = self.NA = 0
        self._NFt = 0
        self._keep = None  # host array kept alive during async staging

    # -- lifecycle
    def close(self):
        if getattr(self, "h", None):
            self.lib.sgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc:
            raise SgpuError(rc, self.lib.sgpu_last_error(self.h).decode())

    def synchronize(self):
        self._ck(self.lib.sgpu_synchronize(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.sgpu_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.sgpu_stream(self.h) or 0)

    def pinned(self, shape, dtype=np.float32) -> PinnedArray:
        return PinnedArray(self.lib, shape, dtype)

    # -- staging
    def stage_frames(self, xyz, repr=REPR_CARTESIAN):
        """xyz: float32 [NF][NA][3] host array (numpy or PinnedArray.array)."""
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        if a.ndim != 3 or a.shape[2] != 3:
            raise SgpuError(1, "stage_frames expects [NF][NA][3]")
        self._keep = a
        self._ck(self.lib.sgpu_stage_frames(self.h, a.ctypes.data, a.shape[0], a.shape[1], int(repr)))
        self.NF, self.NA = a.shape[0], a.shape[1]
        self._NFt = 0

    def stage_frames_device(self, d_ptr: int, NF: int, NA: int, repr=REPR_CARTESIAN):
        self._ck(self.lib.sgpu_stage_frames_device(self.h, C.c_void_p(d_ptr), NF, NA, int(repr)))
        self.NF, self.NA = NF, NA
        self._NFt = 0

    def frames_to_spherical(self):
        self._ck(self.lib.sgpu_frames_to_spherical(self.h))

    def stage_atoms(self, xyz_by_atom):
        """xyz_by_atom: float32 [NA_local][NF][3].  Asynchronous like stage_frames (chunks of atoms on the copy stream; the
        autocorrelation path evaluates them as they land): the array is kept alive here, do not overwrite or free its memory
        before synchronize() or a compute that returns results."""
        a = np.ascontiguousarray(xyz_by_atom, dtype=np.float32)
        if a.ndim != 3 or a.shape[2] != 3:
            raise SgpuError(1, "stage_atoms expects [NA][NF][3]")
        self._keep = a
        self._ck(self.lib.sgpu_stage_atoms(self.h, a.ctypes.data, a.shape[0], a.shape[1]))
        self.NA, self.NF = a.shape[0], a.shape[1]
        self._NFt = 0

    def stage_atoms_device(self, d_ptr: int, NA_local: int, NF: int):
        self._ck(self.lib.sgpu_stage_atoms_device(self.h, C.c_void_p(d_ptr), NA_local, NF))
        self.NA, self.NF = NA_local, NF
        self._NFt = 0

    def stage_atoms_from_frames(self, xyz, nranks=1, rank=0):
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        self._ck(self.lib.sgpu_stage_atoms_from_frames(self.h, a.ctypes.data, a.shape[0], a.shape[1], nranks, rank))
        NA = a.shape[1]
        self.NF = a.shape[0]
        self._NFt = 0
        self.NA = NA // nranks + (1 if rank < NA % nranks else 0)

    def stage_atoms_wave(self, xyz, atom_first, atom_stride, count):
        """atoms atom_first + i*atom_stride, i < count, of frame-major xyz [NF][NA][3] (one wave of a streamed self run)"""
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        self._ck(self.lib.sgpu_stage_atoms_wave(self.h, a.ctypes.data, a.shape[0], a.shape[1], atom_first, atom_stride, count))
        self.NF = a.shape[0]
        self._NFt = 0
        self.NA = count

    def stage_atoms_prefetch(self, xyz_block):
        """queue the H2D copy of a block of atoms (host float32 [count][NF][3], pinned for an asynchronous copy) into the
        back wave buffer; returns at once.  The array must stay alive until the matching swap's first compute has returned."""
        a = xyz_block
        assert a.dtype == np.float32 and a.ndim == 3 and a.shape[2] == 3 and a.flags["C_CONTIGUOUS"]
        self._ck(self.lib.sgpu_stage_atoms_prefetch(self.h, a.ctypes.data, a.shape[0], a.shape[1]))
        self._pending_wave = (a.shape[0], a.shape[1])

    def stage_atoms_swap(self):
        """make the prefetched block the staged atoms (no host synchronisation)"""
        self._ck(self.lib.sgpu_stage_atoms_swap(self.h))
        self.NA, self.NF = self._pending_wave  # the output buffers of finalize / compute are sized from these
        self._NFt = 0

    def device_bytes(self) -> int:
        n = C.c_size_t(0)
        self._ck(self.lib.sgpu_device_bytes(self.h, C.byref(n)))
        return int(n.value)

    def accumulate(self, d_dst: int, d_src: int, n: int):
        self._ck(self.lib.sgpu_accumulate(self.h, C.c_void_p(d_dst), C.c_void_p(d_src), n))

    def set_factors(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        self._ck(self.lib.sgpu_set_factors(self.h, _dp(b), b.size))

    # -- compute
    def _outputs(self):
        return np.zeros(2 * self.timeline_frames), np.zeros(2), np.zeros(2)

    @property
    def timeline_frames(self):
        """frames of the timelines the DSP works on: the window's NF_total if one is set, else the staged NF"""
        # asked of the library, not tracked here: an output buffer sized from a stale shape would be overrun by finalize
        nf, na, nft = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        self._ck(self.lib.sgpu_staged_shape(self.h, C.byref(nf), C.byref(na), C.byref(nft)))
        return int(nft.value) if nft.value else (self._NFt if self._NFt else self.NF)

    @staticmethod
    def _pack(at, af, a2f):
        return at.view(np.complex128).copy(), complex(af[0], af[1]), complex(a2f[0], a2f[1])

    def compute_all_vectors(self, qvecs, dsp="autocorrelate", method="fftw"):
        q = np.ascontiguousarray(qvecs, dtype=np.float64).reshape(-1, 3)
        at, af, a2f = self._outputs()
        self._ck(self.lib.sgpu_compute_all_vectors(self.h, _dp(q), len(q), _dsp(dsp), _method(method), _dp(at), _dp(af),
                                                   _dp(a2f)))
        return self._pack(at, af, a2f)

    def compute_self_vectors(self, qvecs, dsp="autocorrelate", method="fftw"):
        q = np.ascontiguousarray(qvecs, dtype=np.float64).reshape(-1, 3)
        at, af, a2f = self._outputs()
        self._ck(self.lib.sgpu_compute_self_vectors(self.h, _dp(q), len(q), _dsp(dsp), _method(method), _dp(at),
                                                    _dp(af), _dp(a2f)))
        return self._pack(at, af, a2f)

    def compute_mpsphere(self, qlen, moments, dsp="autocorrelate", method="fftw"):
        lm = np.ascontiguousarray(moments, dtype=np.int64).reshape(-1, 2)
        at, af, a2f = self._outputs()
        self._ck(self.lib.sgpu_compute_mpsphere(self.h, float(qlen), lm.ctypes.data_as(C.POINTER(C.c_long)), len(lm),
                                                _dsp(dsp), _method(method), _dp(at), _dp(af), _dp(a2f)))
        return self._pack(at, af, a2f)

    def frames_to_cylindrical(self, axis):
        """staged cartesian frames -> (r, phi, z) in the basis built on `axis` (MPCylinder devices)"""
        ax = np.ascontiguousarray(axis, dtype=np.float64).reshape(3)
        self._ck(self.lib.sgpu_frames_to_cylindrical(self.h, _dp(ax)))

    def compute_mpcylinder(self, q, axis, moments, dsp="autocorrelate", method="fftw"):
        qv = np.ascontiguousarray(q, dtype=np.float64).reshape(3)
        ax = np.ascontiguousarray(axis, dtype=np.float64).reshape(3)
        lm = np.ascontiguousarray(moments, dtype=np.int64).reshape(-1, 2)
        at, af, a2f = self._outputs()
        self._ck(self.lib.sgpu_compute_mpcylinder(self.h, _dp(qv), _dp(ax), lm.ctypes.data_as(C.POINTER(C.c_long)), len(lm),
                                                  _dsp(dsp), _method(method), _dp(at), _dp(af), _dp(a2f)))
        return self._pack(at, af, a2f)

    def compute_mpcylinder_partial(self, q, axis, moments, d_partial: int, dsp="autocorrelate"):
        qv = np.ascontiguousarray(q, dtype=np.float64).reshape(3)
        ax = np.ascontiguousarray(axis, dtype=np.float64).reshape(3)
        lm = np.ascontiguousarray(moments, dtype=np.int64).reshape(-1, 2)
        self._ck(self.lib.sgpu_compute_mpcylinder_partial(self.h, _dp(qv), _dp(ax), lm.ctypes.data_as(C.POINTER(C.c_long)),
                                                          len(lm), _dsp(dsp), C.c_void_p(d_partial)))

    def mpcylinder_amplitudes(self, q, axis, moments, atom_first, atom_count, d_amp: int):
        qv = np.ascontiguousarray(q, dtype=np.float64).reshape(3)
        ax = np.ascontiguousarray(axis, dtype=np.float64).reshape(3)
        lm = np.ascontiguousarray(moments, dtype=np.int64).reshape(-1, 2)
        self._ck(self.lib.sgpu_mpcylinder_amplitudes(self.h, _dp(qv), _dp(ax), lm.ctypes.data_as(C.POINTER(C.c_long)), len(lm),
                                                     atom_first, atom_count, C.c_void_p(d_amp)))

    def set_factors_batch(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        self._ck(self.lib.sgpu_set_factors_batch(self.h, _dp(b), b.shape[0], b.shape[1]))

    def compute_mpsphere_batch(self, qlens, moments, dsp="autocorrelate", method="fftw"):
        """NQ |q| values in one pass -> list of (fqt, fq, fq2)."""
        ql = np.ascontiguousarray(qlens, dtype=np.float64).reshape(-1)
        lm = np.ascontiguousarray(moments, dtype=np.int64).reshape(-1, 2)
        NQ = len(ql)
        at, af, a2f = np.zeros((NQ, 2 * self.NF)), np.zeros((NQ, 2)), np.zeros((NQ, 2))
        self._ck(self.lib.sgpu_compute_mpsphere_batch(self.h, _dp(ql), NQ, lm.ctypes.data_as(C.POINTER(C.c_long)), len(lm),
                                                      _dsp(dsp), _method(method), _dp(at), _dp(af), _dp(a2f)))
        return [self._pack(at[i], af[i], a2f[i]) for i in range(NQ)]

    def mpsphere_amplitudes(self, qlens, moments, atom_first, atom_count, d_amp: int):
        ql = np.ascontiguousarray(qlens, dtype=np.float64).reshape(-1)
        lm = np.ascontiguousarray(moments, dtype=np.int64).reshape(-1, 2)
        self._ck(self.lib.sgpu_mpsphere_amplitudes(self.h, _dp(ql), len(ql), lm.ctypes.data_as(C.POINTER(C.c_long)), len(lm),
                                                   atom_first, atom_count, C.c_void_p(d_amp)))

    def mpsphere_dsp_partial(self, d_amp: int, NQ, NM, d_partials: int, dsp="autocorrelate"):
        self._ck(self.lib.sgpu_mpsphere_dsp_partial(self.h, C.c_void_p(d_amp), NQ, NM, _dsp(dsp), C.c_void_p(d_partials)))

    # -- multi-GPU split
    # -- frame-sharded coherent path: stage a block of frames, place it in the timeline, exchange amplitudes
    def set_frame_window(self, NF_total: int, f_first: int):
        self._ck(self.lib.sgpu_set_frame_window(self.h, int(NF_total), int(f_first)))
        self._NFt = int(NF_total)

    def all_vectors_amplitudes(self, qvecs, d_amp: int):
        """d_amp: device complex [NM][NF_total]; this rank's frame columns are written, all others zeroed"""
        q = np.ascontiguousarray(qvecs, dtype=np.float64).reshape(-1, 3)
        self._ck(self.lib.sgpu_all_vectors_amplitudes(self.h, _dp(q), len(q), C.c_void_p(d_amp)))

    def all_vectors_dsp_partial(self, d_amp: int, m_first: int, m_count: int, d_partial: int, dsp="autocorrelate"):
        self._ck(self.lib.sgpu_all_vectors_dsp_partial(self.h, C.c_void_p(d_amp), int(m_first), int(m_count), _dsp(dsp),
                                                       C.c_void_p(d_partial)))

    # -- |q|-scan coherent path: q_{n,m} = s_n v_m
    def compute_all_vectors_scan(self, v, s, dsp="autocorrelate", method="fftw"):
        """v: directions [NM][3]; s: |q| values [NQ].  Returns (fqt [NQ][NF] complex, fq [NQ] complex, fq2 [NQ] complex)."""
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)
        s = np.ascontiguousarray(s, dtype=np.float64).reshape(-1)
        NF, NQ = self.timeline_frames, len(s)
        at, af, a2f = np.zeros((NQ, 2 * NF)), np.zeros((NQ, 2)), np.zeros((NQ, 2))
        self._ck(self.lib.sgpu_compute_all_vectors_scan(self.h, _dp(v), len(v), _dp(s), NQ, _dsp(dsp), _method(method),
                                                        _dp(at), _dp(af), _dp(a2f)))
        return at.view(np.complex128).copy(), af.view(np.complex128)[:, 0].copy(), a2f.view(np.complex128)[:, 0].copy()

    def compute_all_vectors_scan_partial(self, v, s, d_partials: int, dsp="autocorrelate"):
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)
        s = np.ascontiguousarray(s, dtype=np.float64).reshape(-1)
        self._ck(self.lib.sgpu_compute_all_vectors_scan_partial(self.h, _dp(v), len(v), _dp(s), len(s), _dsp(dsp),
                                                                C.c_void_p(d_partials)))

    # ---- the partition's NCCL communicator inside the library
    def comm_init(self, unique_id: bytes, nranks: int, rank: int):
        """collective: every rank of the partition calls it with the id rank 0 obtained from comm_unique_id()"""
        assert len(unique_id) == 128
        self._ck(self.lib.sgpu_comm_init(self.h, unique_id, nranks, rank))

    def comm_destroy(self):
        self._ck(self.lib.sgpu_comm_destroy(self.h))

    def comm_allreduce(self, d_buf: int, n: int):
        """in-place sum of n doubles in device memory over the ranks, on the compute stream (no host synchronisation)"""
        self._ck(self.lib.sgpu_comm_allreduce(self.h, C.c_void_p(d_buf), n))

    def compute_all_vectors_scan_sharded(self, v, s, d_partials: int, dsp="autocorrelate"):
        """frame-sharded coherent scan: local amplitudes, exchange over NVLink, DSP of this rank's timelines, all-reduce of the
        packed partials; d_partials [len(s)][partial_len] holds the reduced partials on every rank"""
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)
        s = np.ascontiguousarray(s, dtype=np.float64).reshape(-1)
        self._ck(self.lib.sgpu_compute_all_vectors_scan_sharded(self.h, _dp(v), len(v), _dp(s), len(s), _dsp(dsp),
                                                                C.c_void_p(d_partials)))

    def compute_all_vectors_sharded(self, qvecs, d_partial: int, dsp="autocorrelate"):
        q = np.ascontiguousarray(qvecs, dtype=np.float64).reshape(-1, 3)
        self._ck(self.lib.sgpu_compute_all_vectors_sharded(self.h, _dp(q), len(q), _dsp(dsp), C.c_void_p(d_partial)))

    def all_vectors_scan_amplitudes(self, v, s, d_amp: int):
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)
        s = np.ascontiguousarray(s, dtype=np.float64).reshape(-1)
        self._ck(self.lib.sgpu_all_vectors_scan_amplitudes(self.h, _dp(v), len(v), _dp(s), len(s), C.c_void_p(d_amp)))

    def last_scan_plan(self):
        """(plain scan passes, corrected scan passes, |q| values through the general kernel) of the last scan call"""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.sgpu_last_scan_plan(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def partial_len(self, dsp="autocorrelate") -> int:
        n = C.c_size_t()
        self._ck(self.lib.sgpu_partial_len(self.h, _dsp(dsp), C.byref(n)))
        return int(n.value)

    def compute_all_vectors_partial(self, qvecs, d_partial: int, dsp="autocorrelate"):
        q = np.ascontiguousarray(qvecs, dtype=np.float64).reshape(-1, 3)
        self._ck(self.lib.sgpu_compute_all_vectors_partial(self.h, _dp(q), len(q), _dsp(dsp), C.c_void_p(d_partial)))

    def compute_self_vectors_partial(self, qvecs, d_partial: int, dsp="autocorrelate"):
        q = np.ascontiguousarray(qvecs, dtype=np.float64).reshape(-1, 3)
        self._ck(self.lib.sgpu_compute_self_vectors_partial(self.h, _dp(q), len(q), _dsp(dsp), C.c_void_p(d_partial)))

    def compute_mpsphere_partial(self, qlen, moments, d_partial: int, dsp="autocorrelate"):
        lm = np.ascontiguousarray(moments, dtype=np.int64).reshape(-1, 2)
        self._ck(self.lib.sgpu_compute_mpsphere_partial(self.h, float(qlen), lm.ctypes.data_as(C.POINTER(C.c_long)),
                                                        len(lm), _dsp(dsp), C.c_void_p(d_partial)))

    def finalize(self, d_partial: int, scale: float, dsp="autocorrelate", method="fftw"):
        at, af, a2f = self._outputs()
        self._ck(self.lib.sgpu_finalize(self.h, C.c_void_p(d_partial), _dsp(dsp), _method(method), float(scale), _dp(at),
                                        _dp(af), _dp(a2f)))
        return self._pack(at, af, a2f)

    # -- introspection
    def get_amplitudes(self, NM):
        A = np.zeros((NM, self.NF), dtype=np.complex128)
        self._ck(self.lib.sgpu_get_amplitudes(self.h, _dp(A.view(np.float64)), NM, self.NF))
        return A

    def last_amplitude_ms(self) -> float:
        v = C.c_float()
        self._ck(self.lib.sgpu_last_amplitude_ms(self.h, C.byref(v)))
        return float(v.value)

    def last_dsp_ms(self) -> float:
        v = C.c_float()
        self._ck(self.lib.sgpu_last_dsp_ms(self.h, C.byref(v)))
        return float(v.value)

    def timer_start(self):
        self._ck(self.lib.sgpu_timer_start(self.h))

    def timer_stop(self) -> float:
        v = C.c_float()
        self._ck(self.lib.sgpu_timer_stop(self.h, C.byref(v)))
        return float(v.value)

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        self._ck(self.lib.sgpu_measure_fp64_peak(self.h, C.byref(v)))
        return float(v.value)

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        rc = self.lib.sgpu_device_alloc(C.byref(p), nbytes)
        if rc:
            raise SgpuError(rc, "sgpu_device_alloc failed")
        return int(p.value)

    def device_free(self, ptr: int):
        self.lib.sgpu_device_free(C.c_void_p(ptr))

    def memcpy_d2h(self, host_array: np.ndarray, d_ptr: int):
        self._ck(self.lib.sgpu_memcpy_d2h(self.h, host_array.ctypes.data, C.c_void_p(d_ptr), host_array.nbytes))

    def memcpy_h2d(self, d_ptr: int, host_array: np.ndarray):
        a = np.ascontiguousarray(host_array)
        self._ck(self.lib.sgpu_memcpy_h2d(self.h, C.c_void_p(d_ptr), a.ctypes.data, a.nbytes))

    def synth_trajectory(self, d_ptr: int, NF, NA, box, sigma, seed, layout=0, atom0=0, atom_stride=1, NA_out=None,
                         offset=0.0):
        from .synth import step_scale
        NA_out = NA if NA_out is None else NA_out
        self._ck(self.lib.sgpu_synth_trajectory(self.h, C.c_void_p(d_ptr), NF, NA, atom0, atom_stride, NA_out,
                                                C.c_float(box), C.c_float(offset), C.c_float(step_scale(sigma)),
                                                C.c_uint64(seed), layout))


def comm_unique_id() -> bytes:
    """128-byte NCCL unique id (one rank creates it, all ranks pass it to ScatterContext.comm_init)"""
    buf = C.create_string_buffer(128)
    rc = load_library().sgpu_comm_get_unique_id(buf)
    if rc:
        raise SgpuError(rc, "sgpu_comm_get_unique_id failed (libnccl.so.2 not loadable?)")
    return buf.raw
