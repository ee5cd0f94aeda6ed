"""ctypes binding of the CPU oracle (oracle/sassena_oracle.c) plus an independent numpy restatement.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package (sassena_b200/) never imports this module.

PARITY PINNED to the reference's own code (see the header of sassena_oracle.c and DESIGN.md section 2): the reference ships
no golden vectors, but its scatter devices (all / self / multipole sphere / multipole cylinder), stagers, DSP, generators,
database and readers compile where they lie over the shims in oracle/shim* -> oracle/_ref/*.so; the ref_* functions below
run them, tests/golden/ref_*.npz hold their output, and the C port reproduces the devices bit for bit.  Analytic known
answers and the numpy / scipy functions below remain as independent cross-checks (special-function values, FFT).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

DSP_AUTOCORRELATE, DSP_SQUARE, DSP_PLAIN = 0, 1, 2
METHOD_FFTW, METHOD_DIRECT = 0, 1

_DSP = {"autocorrelate": 0, "square": 1, "plain": 2}
_METHOD = {"fftw": 0, "direct": 1}


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc -O3 -DNDEBUG -fopenmp)."""
    src = os.path.join(_HERE, "sassena_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_decomposition_penalty.restype = C.c_size_t
        _lib.orc_scan_unfold.restype = C.c_size_t
        _lib.orc_cylinder_raster_linear.restype = C.c_size_t
        _lib.orc_moments_sphere.restype = C.c_size_t
        _lib.orc_moments_cylinder.restype = C.c_size_t
        _lib.orc_init_subvectors.restype = C.c_size_t
        _lib.orc_sph_bessel.restype = C.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------------------------------------
# decomposition
# ---------------------------------------------------------------------------------------------
def div_assignment(NN, rank, NAF):
    o, s, m = C.c_size_t(), C.c_size_t(), C.c_size_t()
    lib().orc_div_assignment(C.c_size_t(NN), C.c_size_t(rank), C.c_size_t(NAF), C.byref(o), C.byref(s), C.byref(m))
    return o.value, s.value, m.value


def mod_assignment(NN, rank, NAF):
    o, s, m = C.c_size_t(), C.c_size_t(), C.c_size_t()
    lib().orc_mod_assignment(C.c_size_t(NN), C.c_size_t(rank), C.c_size_t(NAF), C.byref(o), C.byref(s), C.byref(m))
    return o.value, s.value, m.value


def decomposition_penalty(NN, NQ, NAF, NNpP, elbytes=1):
    """DecompositionParameters(...).penalty() (decomposition_plan.cpp:28-66)"""
    a, b, c, d = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    return lib().orc_decomposition_penalty(C.c_size_t(NN), C.c_size_t(NQ), C.c_size_t(NAF), C.c_size_t(NNpP), C.c_size_t(elbytes),
                                           C.byref(a), C.byref(b), C.byref(c), C.byref(d))


def decomposition_plan(nn, nq, naf, elbytes, maxbytes, min_utilization=0.95):
    p, ps, pen = C.c_size_t(), C.c_size_t(), C.c_size_t()
    rc = lib().orc_decomposition_plan(C.c_size_t(nn), C.c_size_t(nq), C.c_size_t(naf), C.c_size_t(elbytes),
                                      C.c_size_t(maxbytes), C.c_double(min_utilization), C.byref(p), C.byref(ps),
                                      C.byref(pen))
    return rc, p.value, ps.value, pen.value


# ---------------------------------------------------------------------------------------------
# generators
# ---------------------------------------------------------------------------------------------
def scan_unfold(base, frm, to, points, exponent=1.0):
    out = np.zeros((max(points, 1), 3))
    n = lib().orc_scan_unfold(_p(_f64(base), C.c_double), C.c_double(frm), C.c_double(to), C.c_size_t(points),
                              C.c_double(exponent), _p(out, C.c_double))
    return out[:n]


def qvectors_from_scans(scans):
    """ScatteringVectorsParameters::create_from_scans (parameters.cpp:1125-1189): outer-product sum of <=3 scans.
    scans: list of dicts(base, from, to, points, exponent)."""
    if len(scans) > 3:
        raise ValueError("More than 3 scan definitions are not supported.")
    lists = [scan_unfold(s.get("base", (1, 0, 0)), s.get("from", 0.0), s.get("to", 1.0), s.get("points", 100),
                         s.get("exponent", 1.0)) for s in scans]
    if len(lists) == 1:
        return lists[0].copy()
    if len(lists) == 2:
        return np.array([a + b for a in lists[0] for b in lists[1]])
    return np.array([a + b + c for a in lists[0] for b in lists[1] for c in lists[2]])


def mt19937_stream(seed, n):
    out = np.zeros(n, dtype=np.uint32)
    lib().orc_mt19937_stream(C.c_uint32(seed), C.c_size_t(n), _p(out, C.c_uint32))
    return out


def uniform_on_sphere(seed, dim, count):
    out = np.zeros((count, 3))
    lib().orc_uniform_on_sphere(C.c_uint32(seed), C.c_int(dim), C.c_size_t(count), _p(out, C.c_double))
    return out


def normalize_rows(v):
    v = _f64(v).copy()
    lib().orc_normalize_rows(_p(v, C.c_double), C.c_size_t(len(v)))
    return v


def cylinder_raster_linear(resolution):
    n = lib().orc_cylinder_raster_linear(C.c_size_t(resolution), None)
    out = np.zeros((n, 3))
    lib().orc_cylinder_raster_linear(C.c_size_t(resolution), _p(out, C.c_double))
    return out


def moments_sphere(resolution):
    n = lib().orc_moments_sphere(C.c_long(resolution), None)
    out = np.zeros((n, 2), dtype=np.int64)
    lib().orc_moments_sphere(C.c_long(resolution), _p(out, C.c_long))
    return out


def vector_base(axis):
    out = np.zeros((3, 3))
    lib().orc_vector_base(_p(_f64(axis), C.c_double), _p(out, C.c_double))
    return out


def init_subvectors(kind, q, orient=None, axis=(0, 0, 1)):
    """kind: 'none' | 'sphere' | 'file' | 'cylinder'."""
    t = {"none": 0, "sphere": 1, "file": 1, "cylinder": 2}[kind]
    orient = np.zeros((0, 3)) if orient is None else _f64(orient)
    out = np.zeros((max(len(orient), 1), 3))
    n = lib().orc_init_subvectors(C.c_int(t), _p(_f64(q), C.c_double), _p(orient, C.c_double),
                                  C.c_size_t(len(orient)), _p(_f64(axis), C.c_double), _p(out, C.c_double))
    return out[:n]


def cart_to_spherical(xyz):
    xyz = _f32(xyz)
    out = np.zeros_like(xyz)
    lib().orc_cart_to_spherical(_p(xyz, C.c_float), C.c_size_t(xyz.size // 3), _p(out, C.c_float))
    return out


# ---------------------------------------------------------------------------------------------
# math
# ---------------------------------------------------------------------------------------------
def fft(x, sign=-1):
    a = np.ascontiguousarray(x, dtype=np.complex128).copy()
    lib().orc_fft(_p(a.view(np.float64), C.c_double), C.c_size_t(a.size), C.c_int(sign))
    return a


def auto_correlate_fftw(x):
    NF = len(x)
    a = np.zeros(2 * NF, dtype=np.complex128)
    a[:NF] = x
    lib().orc_auto_correlate_fftw(_p(a.view(np.float64), C.c_double), C.c_size_t(NF))
    return a[:NF].copy()


def auto_correlate_direct(x):
    a = np.ascontiguousarray(x, dtype=np.complex128).copy()
    lib().orc_auto_correlate_direct(_p(a.view(np.float64), C.c_double), C.c_size_t(len(a)))
    return a


def sph_bessel(l, x):
    return lib().orc_sph_bessel(C.c_long(l), C.c_double(x))


def spherical_harmonic(n, m, theta, phi):
    re, im = C.c_double(), C.c_double()
    lib().orc_spherical_harmonic(C.c_long(n), C.c_long(m), C.c_double(theta), C.c_double(phi), C.byref(re),
                                 C.byref(im))
    return complex(re.value, im.value)


# ---------------------------------------------------------------------------------------------
# compute() drivers
# ---------------------------------------------------------------------------------------------
def _res(atfinal, af, a2f):
    return atfinal.view(np.complex128).reshape(-1), complex(af[0], af[1]), complex(a2f[0], a2f[1])


def compute_all_vectors(coords, sfs, qvecs, dsp="autocorrelate", method="fftw", nthreads=1, return_amplitudes=False,
                        framesplit=False):
    """coords float32 [NF][NA][3]; sfs f64 [NA]; qvecs f64 [NM][3] (already expanded subvectors).
    Returns (fqt[NF] complex, fq complex, fq2 complex[, A[NM][NF]])."""
    coords = _f32(coords)
    NF, NA, _ = coords.shape
    sfs, qvecs = _f64(sfs), _f64(qvecs).reshape(-1, 3)
    NM = len(qvecs)
    atfinal = np.zeros(2 * NF)
    af, a2f = np.zeros(2), np.zeros(2)
    if framesplit:
        rc = lib().orc_compute_all_vectors_framesplit(
            _p(coords, C.c_float), C.c_size_t(NF), C.c_size_t(NA), _p(sfs, C.c_double), _p(qvecs, C.c_double),
            C.c_size_t(NM), C.c_int(_DSP[dsp]), C.c_int(_METHOD[method]), C.c_int(nthreads), _p(atfinal, C.c_double),
            _p(af, C.c_double), _p(a2f, C.c_double))
        if rc:
            raise RuntimeError("oracle: DSP type/method not understood")
        return _res(atfinal, af, a2f)
    at_out = np.zeros((NM, NF), dtype=np.complex128) if return_amplitudes else None
    rc = lib().orc_compute_all_vectors(
        _p(coords, C.c_float), C.c_size_t(NF), C.c_size_t(NA), _p(sfs, C.c_double), _p(qvecs, C.c_double),
        C.c_size_t(NM), C.c_int(_DSP[dsp]), C.c_int(_METHOD[method]), C.c_int(nthreads), _p(atfinal, C.c_double),
        _p(af, C.c_double), _p(a2f, C.c_double),
        _p(at_out.view(np.float64), C.c_double) if return_amplitudes else None)
    if rc:
        raise RuntimeError("oracle: DSP type/method not understood")
    r = _res(atfinal, af, a2f)
    return r + (at_out,) if return_amplitudes else r


def compute_self_vectors(coords_by_atom, sfs_local, qvecs, dsp="autocorrelate", method="fftw", nthreads=1):
    """coords float32 [NA_local][NF][3]; sfs_local f64 [NA_local]."""
    coords = _f32(coords_by_atom)
    NA, NF, _ = coords.shape
    sfs, qvecs = _f64(sfs_local), _f64(qvecs).reshape(-1, 3)
    atfinal = np.zeros(2 * NF)
    af, a2f = np.zeros(2), np.zeros(2)
    rc = lib().orc_compute_self_vectors(
        _p(coords, C.c_float), C.c_size_t(NA), C.c_size_t(NF), _p(sfs, C.c_double), _p(qvecs, C.c_double),
        C.c_size_t(len(qvecs)), C.c_int(_DSP[dsp]), C.c_int(_METHOD[method]), C.c_int(nthreads),
        _p(atfinal, C.c_double), _p(af, C.c_double), _p(a2f, C.c_double))
    if rc:
        raise RuntimeError("oracle: DSP type/method not understood")
    return _res(atfinal, af, a2f)


def compute_mpsphere(coords_sph, sfs, ql, moments, dsp="autocorrelate", method="fftw", nthreads=1,
                     return_amplitudes=False):
    """coords_sph float32 [NF][NA][3] holding (r, phi, theta); moments int [NM][2] (l, m)."""
    coords = _f32(coords_sph)
    NF, NA, _ = coords.shape
    sfs = _f64(sfs)
    mom = np.ascontiguousarray(moments, dtype=np.int64).reshape(-1, 2)
    NM = len(mom)
    atfinal = np.zeros(2 * NF)
    af, a2f = np.zeros(2), np.zeros(2)
    at_out = np.zeros((NM, NF), dtype=np.complex128) if return_amplitudes else None
    rc = lib().orc_compute_mpsphere(
        _p(coords, C.c_float), C.c_size_t(NF), C.c_size_t(NA), _p(sfs, C.c_double), C.c_double(ql),
        _p(mom, C.c_long), C.c_size_t(NM), C.c_int(_DSP[dsp]), C.c_int(_METHOD[method]), C.c_int(nthreads),
        _p(atfinal, C.c_double), _p(af, C.c_double), _p(a2f, C.c_double),
        _p(at_out.view(np.float64), C.c_double) if return_amplitudes else None)
    if rc == 2:
        raise RuntimeError("oracle: Combination of Major and minor moment not allowed")
    if rc:
        raise RuntimeError("oracle: DSP type/method not understood")
    r = _res(atfinal, af, a2f)
    return r + (at_out,) if return_amplitudes else r


def moments_cylinder(resolution):
    """(0,0) then (l, 0..3) for l = 1..resolution (parameters.cpp:1062-1070)"""
    n = lib().orc_moments_cylinder(C.c_long(resolution), None)
    out = np.zeros((n, 2), dtype=np.int64)
    lib().orc_moments_cylinder(C.c_long(resolution), _p(out, C.c_long))
    return out


def cart_to_cylindrical(xyz, axis):
    """float32 [...][3] cartesian -> (r, phi, z) in the basis built on axis, narrowed to float32 like the stager"""
    a = _f32(xyz)
    out = np.empty_like(a)
    ax = _f64(np.asarray(axis, dtype=np.float64))
    lib().orc_cart_to_cylindrical(_p(a, C.c_float), C.c_size_t(a.size // 3), _p(ax, C.c_double), _p(out, C.c_float))
    return out


def compute_mpcylinder(coords_cyl, sfs, q, axis, moments, dsp="autocorrelate", method="fftw", nthreads=1,
                       return_amplitudes=False):
    """coords_cyl float32 [NF][NA][3] holding (r, phi, z); q the q-vector; moments int [NM][2] (l, m in 0..3)."""
    coords = _f32(coords_cyl)
    NF, NA, _ = coords.shape
    sfs = _f64(sfs)
    qv = _f64(np.asarray(q, dtype=np.float64))
    ax = _f64(np.asarray(axis, dtype=np.float64))
    mom = np.ascontiguousarray(moments, dtype=np.int64).reshape(-1, 2)
    NM = len(mom)
    atfinal = np.zeros(2 * NF)
    af, a2f = np.zeros(2), np.zeros(2)
    at_out = np.zeros((NM, NF), dtype=np.complex128) if return_amplitudes else None
    rc = lib().orc_compute_mpcylinder(
        _p(coords, C.c_float), C.c_size_t(NF), C.c_size_t(NA), _p(sfs, C.c_double), _p(qv, C.c_double), _p(ax, C.c_double),
        _p(mom, C.c_long), C.c_size_t(NM), C.c_int(_DSP[dsp]), C.c_int(_METHOD[method]), C.c_int(nthreads),
        _p(atfinal, C.c_double), _p(af, C.c_double), _p(a2f, C.c_double),
        _p(at_out.view(np.float64), C.c_double) if return_amplitudes else None)
    if rc == 2:
        raise RuntimeError("oracle: multipole moment not allowed for the cylinder")
    if rc:
        raise RuntimeError("oracle: DSP type/method not understood")
    r = _res(atfinal, af, a2f)
    return r + (at_out,) if return_amplitudes else r


def np_mpcylinder_amplitudes(coords_cyl, sfs, q, axis, moments):
    """independent restatement of multipole_scatter_device.cpp:905-985 with scipy.special.jv; A [NM][NF] complex"""
    from scipy.special import jv
    base = vector_base(axis)
    qp = base @ np.asarray(q, dtype=np.float64)
    qr = np.hypot(qp[0], qp[1])
    # CylinderCoor3D's azimuth (coor3d.cpp:119-129): the quadrant offsets are FLOAT pi / pi/2 (sign() returns float)
    if qp[0] != 0.0:
        qphi = np.arctan(qp[1] / qp[0])
        if qp[0] < 0.0:
            qphi += float(np.float32(np.pi)) * (-1.0 if qp[1] < 0.0 else 1.0)
    elif qp[1] != 0.0:
        qphi = float(np.float32(np.pi / 2)) * (-1.0 if qp[1] < 0.0 else 1.0)
    else:
        qphi = 0.0
    if qphi < 0:
        qphi += 2 * np.pi
    qz = qp[2]
    c = np.asarray(coords_cyl, dtype=np.float64)
    r, phi, z = c[..., 0], c[..., 1], c[..., 2]
    w = np.asarray(sfs)[None, :] * np.exp(1j * np.abs(z * qz))
    p, psi = r * qr, phi - qphi
    out = []
    for l, m in np.asarray(moments).reshape(-1, 2):
        if l == 0:
            f = jv(0, p)
        elif m < 2:
            n = 2 * l
            f = np.sqrt(0.5) * 2.0 * (-1.0) ** l * jv(n, p) * (np.cos(n * psi) if m == 0 else np.sin(n * psi))
        else:
            n = 2 * l - 1
            f = 1j * np.sqrt(0.5) * 2.0 * (-1.0) ** (l - 1) * jv(n, p) * (np.cos(n * psi) if m == 2 else np.sin(n * psi))
        out.append(np.sqrt(2 * np.pi) * np.sum(w * f, axis=1))
    return np.array(out)


def max_threads():
    return lib().orc_max_threads()


# ---------------------------------------------------------------------------------------------
# Independent numpy restatement (used to pin the C port; numpy.fft replaces FFTW3)
# ---------------------------------------------------------------------------------------------
def np_amplitudes_all(coords, sfs, qvecs):
    """A[m][f] = sum_j b_j exp(i q_m . r_j(f))  (all_vectors_scatter_device.cpp:418-438), float64 math."""
    c = np.asarray(coords, dtype=np.float32).astype(np.float64)
    ph = np.einsum("fac,mc->mfa", c, _f64(qvecs))
    return (np.cos(ph) * sfs).sum(-1) + 1j * (np.sin(ph) * sfs).sum(-1)


def np_autocorrelate(a, method="fftw"):
    """smath.cpp:141-156 with numpy.fft, or :51-76 as an explicit O(N^2) sum."""
    NF = len(a)
    if method == "fftw":
        X = np.fft.fft(a, 2 * NF)
        c = np.fft.ifft(np.abs(X) ** 2) * (2 * NF)
        return c[:NF] / (2 * NF * (NF - np.arange(NF)))
    out = np.zeros(NF, dtype=np.complex128)
    for tau in range(NF):
        out[tau] = np.sum(a[: NF - tau] * np.conj(a[tau:])) / (NF - tau)
    return out


def np_dsp_store(A, dsp="autocorrelate", method="fftw", norm=None):
    """dsp()+store()+final scale for a stack of timelines A[M][NF] (all_vectors_scatter_device.cpp:209-236,355-360)."""
    M, NF = A.shape
    if dsp == "autocorrelate":
        T = np.array([np_autocorrelate(a, method) for a in A])
    elif dsp == "square":
        T = (np.abs(A) ** 2).astype(np.complex128)
    else:
        T = A.astype(np.complex128)
    a = T.mean(axis=1)
    norm = (1.0 / M) if norm is None else norm
    return T.sum(0) * norm, a.sum() * norm, (a * np.conj(a)).sum() * norm


# ---------------------------------------------------------------------------------------------
# oracle/_ref/libsmath_ref.so: the REFERENCE's own smath.cpp / coor3d.cpp / assignment.cpp compiled where they lie
# (make -C oracle ref; shims under oracle/shim stand in for the headers they merely include).  Checker of the checker.
# ---------------------------------------------------------------------------------------------
_REF_SMATH = os.path.join(_HERE, "_ref", "libsmath_ref.so")
_ref = None


def have_ref_smath():
    return os.path.exists(_REF_SMATH)


def ref_lib():
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF_SMATH)
    return _ref


def ref_auto_correlate_direct(x, vector_overload=False):
    a = np.ascontiguousarray(x, dtype=np.complex128).copy()
    f = ref_lib().ref_auto_correlate_direct_vec if vector_overload else ref_lib().ref_auto_correlate_direct
    f(_p(a.view(np.float64), C.c_double), C.c_size_t(len(a)))
    return a


def ref_auto_correlate_fftw(x):
    NF = len(x)
    a = np.zeros(2 * NF, dtype=np.complex128)
    a[:NF] = x
    ref_lib().ref_auto_correlate_fftw(_p(a.view(np.float64), C.c_double), C.c_size_t(NF))
    return a[:NF].copy()


def ref_square_elements(x):
    a = np.ascontiguousarray(x, dtype=np.complex128).copy()
    ref_lib().ref_square_elements(_p(a.view(np.float64), C.c_double), C.c_size_t(len(a)))
    return a


def ref_cart_to_spherical(xyz):
    """double [n][3] -> (r, phi, theta) by SphericalCoor3D(CartesianCoor3D), in double (the stager narrows afterwards)"""
    a = _f64(xyz).reshape(-1, 3)
    out = np.empty_like(a)
    for i in range(len(a)):
        ref_lib().ref_cart_to_spherical(_p(a[i], C.c_double), _p(out[i], C.c_double))
    return out


def ref_cart_to_cylindrical(xyz, axis):
    a = _f64(xyz).reshape(-1, 3)
    ax = _f64(np.asarray(axis, dtype=np.float64))
    out = np.empty_like(a)
    for i in range(len(a)):
        ref_lib().ref_cart_to_cylindrical(_p(a[i], C.c_double), _p(ax, C.c_double), _p(out[i], C.c_double))
    return out


def ref_vector_base(axis):
    out = np.zeros((3, 3))
    ref_lib().ref_vector_base(_p(_f64(np.asarray(axis, dtype=np.float64)), C.c_double), _p(out, C.c_double))
    return out


def ref_assignment(mod, NN, rank, NAF):
    """(offset, size, max, indices) of the reference's ModAssignment (mod=True) / DivAssignment"""
    o, s, m = C.c_size_t(), C.c_size_t(), C.c_size_t()
    idx = np.zeros(max(NAF, 1), dtype=np.uintp)
    ref_lib().ref_assignment(C.c_int(1 if mod else 0), C.c_size_t(NN), C.c_size_t(rank), C.c_size_t(NAF), C.byref(o), C.byref(s),
                             C.byref(m), idx.ctypes.data_as(C.POINTER(C.c_size_t)), C.c_size_t(len(idx)))
    return o.value, s.value, m.value, idx[:s.value].astype(np.int64)


def ref_decomposition_penalty(NN, NQ, NAF, NNpP):
    f = ref_lib().ref_decomposition_penalty
    f.restype = C.c_size_t
    return f(C.c_size_t(NN), C.c_size_t(NQ), C.c_size_t(NAF), C.c_size_t(NNpP))


def ref_decomposition_plan(nn, nq, naf, elbytes, maxbytes, automatic=True, manual_size=1, utilization=0.95):
    """(partitions, partitionsize, penalty, colors) out of the reference's DecompositionPlan.  ONLY for inputs with a valid
    plan -- the reference fails with a bare `throw;`, which terminates the process."""
    p, ps, pen = C.c_size_t(), C.c_size_t(), C.c_size_t()
    col = np.zeros(nn, dtype=np.uintp)
    ref_lib().ref_decomposition_plan(C.c_size_t(nn), C.c_size_t(nq), C.c_size_t(naf), C.c_size_t(elbytes), C.c_size_t(maxbytes),
                                     C.c_int(1 if automatic else 0), C.c_size_t(manual_size), C.c_double(utilization),
                                     C.byref(p), C.byref(ps), C.byref(pen), col.ctypes.data_as(C.POINTER(C.c_size_t)))
    return p.value, ps.value, pen.value, col.astype(np.int64)


def ref_dcd_write(path, xyz, split=None):
    """the reference's DCDCoordinateWriter (coordinate_writer.cpp): float [blocks][entries][3] -> DCD file; `split`: written
    in two pieces (blocks [split, n) first), the way two ranks of a partition write it"""
    a = np.ascontiguousarray(xyz, dtype=np.float32)
    p = a.ctypes.data_as(C.POINTER(C.c_float))
    if split is None:
        ref_lib().ref_dcd_write(str(path).encode(), p, C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1]))
    else:
        ref_lib().ref_dcd_write_split(str(path).encode(), p, C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1]), C.c_size_t(split))


_WALKERS = {"linear": 0, "fixed": 1, "oscillation": 2, "randomwalk": 3, "brownian": 4, "localbrownian": 5, "rotationalbrownian": 6}


def ref_motion_transforms(kind, n, displace=0.0, frequency=0.001, radius=0.0, seed=0, sampling=1, direction=(1, 0, 0)):
    """the reference's own motion walkers (src/sample/motion_walker.cpp over the uBLAS / Boost.Random shims): [n][4][4]"""
    d = _f64(np.asarray(direction, dtype=np.float64))
    out = np.zeros((n, 4, 4))
    rc = ref_lib().ref_motion_transforms(C.c_int(_WALKERS[kind]), C.c_double(displace), C.c_double(frequency), C.c_double(radius),
                                         C.c_ulong(seed), C.c_long(sampling), _p(d, C.c_double), C.c_size_t(n), _p(out, C.c_double))
    assert rc == 0
    return out


def ref_scatter_run(kind, frames, b, qvectors, orient=None, vectors_type="file", axis=(0, 0, 1), dsp="autocorrelate", method="fftw",
                    threads=1):
    """the REFERENCE's own AllVectorsScatterDevice ("all") / SelfVectorsScatterDevice ("self") for one rank (oracle/_ref build over
    the shims).  frames float32 [NF][NA][3]; orient: unit vectors [NM][3] or None.  Returns (qvectors written, fqt, fq, fq2)."""
    fr = _f32(frames)
    NF, NA, _ = fr.shape
    bb = _f64(b)
    qv = _f64(qvectors).reshape(-1, 3)
    NQ = len(qv)
    ori = np.zeros((0, 3)) if orient is None else _f64(orient).reshape(-1, 3)
    ax = _f64(np.asarray(axis, dtype=np.float64))
    fqt, fq, fq2, qout = np.zeros((NQ, NF, 2)), np.zeros((NQ, 2)), np.zeros((NQ, 2)), np.zeros((NQ, 3))
    f = ref_lib().ref_scatter_run
    f.restype = C.c_size_t
    n = f(C.c_int(0 if kind == "all" else 1), _p(fr, C.c_float), C.c_size_t(NF), C.c_size_t(NA), _p(bb, C.c_double),
          _p(qv, C.c_double), C.c_size_t(NQ), vectors_type.encode(), _p(ori, C.c_double), C.c_size_t(len(ori)), _p(ax, C.c_double),
          dsp.encode(), method.encode(), C.c_size_t(threads), _p(fqt, C.c_double), _p(fq, C.c_double), _p(fq2, C.c_double),
          _p(qout, C.c_double))
    assert n == NQ
    return qout, fqt[..., 0] + 1j * fqt[..., 1], fq[:, 0] + 1j * fq[:, 1], fq2[:, 0] + 1j * fq2[:, 1]


def ref_scatter_run_ranks(kind, nranks, frames, b, qvectors, orient=None, vectors_type="file", axis=(0, 0, 1), dsp="autocorrelate",
                          method="fftw", threads=1):
    """the REFERENCE's own scatter device on `nranks` ranks of one partition (threads over the shared-memory communicator shim):
    its frame decomposition with all_to_all + alignpad ("all"), its atom decomposition with the staged transposition of
    DataStagerByAtom ("self") and the reductions to partition rank 0 run as written.  Returns like ref_scatter_run."""
    fr = _f32(frames)
    NF, NA, _ = fr.shape
    bb = _f64(b)
    qv = _f64(qvectors).reshape(-1, 3)
    NQ = len(qv)
    ori = np.zeros((0, 3)) if orient is None else _f64(orient).reshape(-1, 3)
    ax = _f64(np.asarray(axis, dtype=np.float64))
    fqt, fq, fq2, qout = np.zeros((NQ, NF, 2)), np.zeros((NQ, 2)), np.zeros((NQ, 2)), np.zeros((NQ, 3))
    f = ref_lib().ref_scatter_run_ranks
    f.restype = C.c_size_t
    n = f(C.c_int(0 if kind == "all" else 1), C.c_int(nranks), _p(fr, C.c_float), C.c_size_t(NF), C.c_size_t(NA), _p(bb, C.c_double),
          _p(qv, C.c_double), C.c_size_t(NQ), vectors_type.encode(), _p(ori, C.c_double), C.c_size_t(len(ori)), _p(ax, C.c_double),
          dsp.encode(), method.encode(), C.c_size_t(threads), _p(fqt, C.c_double), _p(fq, C.c_double), _p(fq2, C.c_double),
          _p(qout, C.c_double))
    assert n == NQ, n
    return qout, fqt[..., 0] + 1j * fqt[..., 1], fq[:, 0] + 1j * fq[:, 1], fq2[:, 0] + 1j * fq2[:, 1]


def ref_multipole_run(kind, frames, b, qvectors, moments, axis=(0, 0, 1), dsp="autocorrelate", method="fftw", threads=1):
    """the REFERENCE's own MPSphereScatterDevice ("sphere") / MPCylinderScatterDevice ("cylinder") for one rank (oracle/_ref build
    over the shims; Boost.Math's sph_bessel / spherical_harmonic / cyl_bessel_j served by the oracle's restatements).  frames
    float32 [NF][NA][3] CARTESIAN (the device asks its sample for the spherical / cylindrical representation); moments int
    [NMOM][2].  Returns (qvectors written, fqt, fq, fq2)."""
    fr = _f32(frames)
    NF, NA, _ = fr.shape
    bb = _f64(b)
    qv = _f64(qvectors).reshape(-1, 3)
    NQ = len(qv)
    mom = np.ascontiguousarray(moments, dtype=np.int64).reshape(-1, 2)
    ax = _f64(np.asarray(axis, dtype=np.float64))
    fqt, fq, fq2, qout = np.zeros((NQ, NF, 2)), np.zeros((NQ, 2)), np.zeros((NQ, 2)), np.zeros((NQ, 3))
    f = ref_lib().ref_multipole_run
    f.restype = C.c_size_t
    n = f(C.c_int(2 if kind == "sphere" else 3), _p(fr, C.c_float), C.c_size_t(NF), C.c_size_t(NA), _p(bb, C.c_double),
          _p(qv, C.c_double), C.c_size_t(NQ), _p(mom, C.c_long), C.c_size_t(len(mom)), _p(ax, C.c_double), dsp.encode(),
          method.encode(), C.c_size_t(threads), _p(fqt, C.c_double), _p(fq, C.c_double), _p(fq2, C.c_double), _p(qout, C.c_double))
    assert n == NQ
    return qout, fqt[..., 0] + 1j * fqt[..., 1], fq[:, 0] + 1j * fq[:, 1], fq2[:, 0] + 1j * fq2[:, 1]


# the reference's own generators (src/control/parameters.cpp:930-1189 in oracle/_ref/libparams_ref.so; oracle/ref_params_wrap.cpp)
_REF_PARAMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libparams_ref.so")
_ref_params = None


def have_ref_params():
    return os.path.exists(_REF_PARAMS)


def ref_params_lib():
    global _ref_params
    if _ref_params is None:
        _ref_params = C.CDLL(_REF_PARAMS)
    return _ref_params


def ref_orientation_vectors(type, algorithm="boost_uniform_on_sphere", resolution=100, seed=0, filepath=""):
    f = ref_params_lib().ref_orientation_vectors
    f.restype = C.c_size_t
    args = (type.encode(), algorithm.encode(), C.c_size_t(resolution), C.c_ulong(seed), filepath.encode())
    n = f(*args, None, C.c_size_t(0))
    out = np.zeros((n, 3))
    f(*args, _p(out, C.c_double), C.c_size_t(n))
    return out


def ref_multipole_moments(multipole_type, resolution=20, type="resolution", filepath=""):
    f = ref_params_lib().ref_multipole_moments
    f.restype = C.c_size_t
    args = (multipole_type.encode(), type.encode(), C.c_long(resolution), filepath.encode())
    n = f(*args, None, C.c_size_t(0))
    out = np.zeros((n, 2), dtype=np.int64)
    f(*args, _p(out, C.c_long), C.c_size_t(n))
    return out


def ref_scan_vectors(scans):
    """scans: list of dicts (base, from, to, points, exponent) as qvectors_from_scans takes them"""
    n = len(scans)
    fr = _f64(np.array([s["from"] for s in scans], dtype=np.float64))
    to = _f64(np.array([s["to"] for s in scans], dtype=np.float64))
    pts = np.ascontiguousarray([s.get("points", 100) for s in scans], dtype=np.uint64)
    ex = _f64(np.array([s.get("exponent", 1.0) for s in scans], dtype=np.float64))
    base = _f64(np.array([s.get("base", (1, 0, 0)) for s in scans], dtype=np.float64).reshape(-1, 3))
    f = ref_params_lib().ref_scan_vectors
    f.restype = C.c_size_t
    args = (C.c_size_t(n), _p(fr, C.c_double), _p(to, C.c_double), pts.ctypes.data_as(C.POINTER(C.c_size_t)), _p(ex, C.c_double),
            _p(base, C.c_double))
    cnt = f(*args, None, C.c_size_t(0))
    out = np.zeros((cnt, 3))
    f(*args, _p(out, C.c_double), C.c_size_t(cnt))
    return out


# the reference's own structure / selection readers (atoms.cpp, atomselection_reader.cpp, atomselection.cpp; ref_sample_wrap.cpp)
def ref_sample_name_reg(label, regexp):
    ref_params_lib().ref_sample_name_reg(label.encode(), regexp.encode())


def ref_atoms_labels(pdbfile):
    """database label of every ATOM record of a PDB structure file (Atoms::add, atoms.cpp:64-86)"""
    f = ref_params_lib().ref_atoms_labels
    f.restype = C.c_size_t
    buf = C.create_string_buffer(1 << 20)
    n = f(pdbfile.encode(), buf, C.c_size_t(1 << 20))
    labels = buf.value.decode().split("\n")[:-1]
    assert len(labels) == n
    return labels


def _ref_sel(fn, *args):
    fn.restype = C.c_size_t
    n = fn(*args, None, C.c_size_t(0))
    if n == C.c_size_t(-1).value:
        return None
    out = np.zeros(n, dtype=np.uint64)
    fn(*args, out.ctypes.data_as(C.POINTER(C.c_size_t)), C.c_size_t(n))
    return out.astype(np.int64)


def ref_select_pdb(file, selector, expression):
    return _ref_sel(ref_params_lib().ref_select_pdb, file.encode(), selector.encode(), expression.encode())


def ref_select_ndx(file, selector, expression, group):
    """one group of the ndx file after the expression filter, or None when the reference did not create it"""
    return _ref_sel(ref_params_lib().ref_select_ndx, file.encode(), selector.encode(), expression.encode(), group.encode())


def ref_select_range(first, last):
    return _ref_sel(ref_params_lib().ref_select_range, C.c_size_t(first), C.c_size_t(last))



# the reference's own rotational fit (src/sample/center_of_mass.cpp:53-157; oracle/ref_cofm_wrap.cpp).  Its SVD is LAPACK's dgesvd,
# reached through the Boost.Bindings call: the shim binds the dgesvd of the OpenBLAS that scipy bundles (a real LAPACK).
def _lapack_lib():
    import glob
    import scipy
    c = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so"))
    return c[0] if c else None


def have_ref_fit():
    return have_ref_params() and _lapack_lib() is not None and hasattr(ref_params_lib(), "ref_fit")


def ref_mass_reg(label, mass):
    ref_params_lib().ref_mass_reg(label.encode(), C.c_double(mass))


def ref_fit(pdbfile, xyz, xyz_ref, sel_ref, sel_manip):
    """Fit(atoms, cs, all, manip, cs_ref, ref): the frame xyz [natoms][3] fitted onto the sel_ref atoms of the frame xyz_ref
    (mass-weighted least-squares rotation + translation onto the reference's centre of mass), the sel_manip atoms moved.
    Returns (fitted coordinates [natoms][3], centre of mass of the sel_ref atoms of xyz before the fit)."""
    os.environ["ORACLE_LAPACK_LIB"] = _lapack_lib()
    a = _f64(np.asarray(xyz, dtype=np.float64).reshape(-1, 3))
    r = _f64(np.asarray(xyz_ref, dtype=np.float64).reshape(-1, 3))
    sr = np.ascontiguousarray(sel_ref, dtype=np.uint64)
    sm = np.ascontiguousarray(sel_manip, dtype=np.uint64)
    out = np.zeros_like(a)
    com = np.zeros(3)
    ref_params_lib().ref_fit(str(pdbfile).encode(), _p(a, C.c_double), _p(r, C.c_double), C.c_size_t(len(a)),
                             sr.ctypes.data_as(C.POINTER(C.c_size_t)), C.c_size_t(len(sr)),
                             sm.ctypes.data_as(C.POINTER(C.c_size_t)), C.c_size_t(len(sm)), _p(out, C.c_double), _p(com, C.c_double))
    return out, com


def ref_frames_read(format, file, first=0, last=0, last_set=False, stride=1):
    """the reference's own DCD / PDB / XTC / TRR frameset (frames.cpp; generate_index + trim_index + read_frame) -> float64
    [nframes][natoms][3]"""
    f = ref_params_lib().ref_frames_read
    f.restype = C.c_size_t
    na = C.c_size_t(0)
    args = (format.encode(), str(file).encode(), C.c_size_t(first), C.c_size_t(last), C.c_int(1 if last_set else 0), C.c_size_t(stride))
    n = f(*args, C.byref(na), None, C.c_size_t(0))
    assert n != C.c_size_t(-1).value, "unknown format"
    out = np.zeros((n, na.value, 3))
    f(*args, C.byref(na), _p(out, C.c_double), C.c_size_t(n))
    return out


def ref_coordinate_set(xyz, sel=None, trans=(0, 0, 0), repr="cartesian", axis=(0, 0, 1)):
    """the reference's own CartesianCoordinateSet(frame, selection) -> translate -> Spherical / CylindricalCoordinateSet
    (coordinate_set.cpp); xyz [natoms][3] -> double [nsel][3]"""
    a = _f64(np.asarray(xyz, dtype=np.float64).reshape(-1, 3))
    idx = np.ascontiguousarray(np.arange(len(a)) if sel is None else sel, dtype=np.uint64)
    out = np.zeros((len(idx), 3))
    ref_params_lib().ref_coordinate_set(_p(a, C.c_double), C.c_size_t(len(a)), idx.ctypes.data_as(C.POINTER(C.c_size_t)),
                                        C.c_size_t(len(idx)), _p(_f64(np.asarray(trans, dtype=np.float64)), C.c_double),
                                        C.c_int({"cartesian": 10, "spherical": 20, "cylindrical": 30}[repr]),
                                        _p(_f64(np.asarray(axis, dtype=np.float64)), C.c_double), _p(out, C.c_double))
    return out


# the reference's own Database tables (src/control/database.cpp in oracle/_ref; entry points oracle/ref_db_wrap.cpp)
REF_DB_TABLES = {"sizes": 0, "exclusionfactors": 1, "scatterfactors": 2}


def ref_db_reg(table, ID, constants, function_type):
    c = _f64(np.asarray(constants, dtype=np.float64))
    ref_lib().ref_db_reg(C.c_int(REF_DB_TABLES[table]), C.c_size_t(ID), _p(c, C.c_double), C.c_size_t(len(c)), C.c_size_t(function_type))


def ref_db_volume(ID):
    f = ref_lib().ref_db_volume
    f.restype = C.c_double
    return f(C.c_size_t(ID))


def ref_db_exclusion(ID, effvolume, q):
    f = ref_lib().ref_db_exclusion
    f.restype = C.c_double
    return f(C.c_size_t(ID), C.c_double(effvolume), C.c_double(q))


def ref_db_sfactor(ID, q):
    f = ref_lib().ref_db_sfactor
    f.restype = C.c_double
    return f(C.c_size_t(ID), C.c_double(q))


def ref_db_effective(ID, q, kappa, background_sl):
    """ScatterFactors::update for one atom (scatter_factors.cpp:56-78) over the reference's Database"""
    f = ref_lib().ref_db_effective
    f.restype = C.c_double
    return f(C.c_size_t(ID), C.c_double(q), C.c_double(kappa), C.c_double(background_sl))


def ref_db_name_reg(label, regexp):
    ref_lib().ref_db_name_reg(label.encode(), regexp.encode())


def ref_db_name_get(testlabel):
    """element label of a PDB atom name the database knows (unknown names end in the reference's bare `throw;`)"""
    buf = C.create_string_buffer(256)
    ref_lib().ref_db_name_get(testlabel.encode(), buf, C.c_size_t(256))
    return buf.value.decode()


def ref_timer_seconds(key):
    """seconds the last ref_scatter_run spent under one of the reference's timer keys ("sd:stage", "sd:runner", "sd:compute", ...)"""
    f = ref_lib().ref_timer_seconds
    f.restype = C.c_double
    return f(key.encode())
