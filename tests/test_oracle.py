"""Pins the CPU oracle (oracle/sassena_oracle.c).  The reference ships no golden vectors or asserting tests for this path
(SURVEY 8c); the pins are the reference's own code built in oracle/_ref (fixtures tests/golden/ref_*.npz written by it:
smath / coor3d / assignment here, the scatter devices in tests/test_reference_devices.py, the generators in
tests/test_reference_params.py), plus independent numpy / scipy restatements and the analytic known answers KA1-KA5, KA7
derived from the cited formulas."""
import os

import numpy as np
import pytest
import scipy.special as sp

from sassena_b200 import synth
from util import rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("n", [1, 2, 3, 8, 30, 97, 200, 202, 1000, 2000, 2018, 4096, 6006])
def test_fft_matches_numpy(oracle, n):
    """own mixed-radix / Bluestein FFT (stands in for FFTW3) vs numpy.fft, both directions"""
    rng = np.random.default_rng(n)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    assert rel_err(oracle.fft(x, -1), np.fft.fft(x)) < 1e-13
    assert rel_err(oracle.fft(x, +1), np.fft.ifft(x) * n) < 1e-13


@pytest.mark.parametrize("NF", [1, 2, 7, 100, 257])
def test_autocorrelation_forms(oracle, NF):
    """smath.cpp:141-156 (fftw) vs numpy; smath.cpp:51-76 (direct) vs an explicit sum; KA4: direct == conj(fftw)"""
    rng = np.random.default_rng(NF)
    a = rng.normal(size=NF) + 1j * rng.normal(size=NF)
    f = oracle.auto_correlate_fftw(a)
    d = oracle.auto_correlate_direct(a)
    assert rel_err(f, oracle.np_autocorrelate(a, "fftw")) < 1e-13
    assert rel_err(d, oracle.np_autocorrelate(a, "direct")) < 1e-13
    assert rel_err(d, np.conj(f)) < 1e-12
    # definition check of the fftw form: c[tau] = sum_k conj(a_k) a_{k+tau} / (NF - tau)
    ref = np.array([np.sum(np.conj(a[:NF - t]) * a[t:]) / (NF - t) for t in range(NF)])
    assert rel_err(f, ref) < 1e-13


def test_special_functions_match_scipy(oracle):
    """sph_bessel / spherical_harmonic stand-ins for Boost.Math vs scipy.special (same conventions as Boost:
    theta polar, phi azimuth, Condon-Shortley)"""
    for l in range(0, 31):
        for x in np.concatenate([np.logspace(-6, 2.5, 120), [0.0, 1.0, float(l), l + 0.5]]):
            r = sp.spherical_jn(l, x)
            v = oracle.sph_bessel(l, x)
            assert abs(v - r) <= 1e-12 * max(abs(r), 1e-3 / max(x, 1.0)), (l, x, v, r)
    for n in range(0, 21):
        for m in range(-n, n + 1):
            for th, ph in [(0.0, 0.0), (0.3, 1.0), (np.pi / 2, 4.0), (2.9, 6.0), (np.pi, 0.5)]:
                assert abs(oracle.spherical_harmonic(n, m, th, ph) - sp.sph_harm_y(n, m, th, ph)) < 1e-13
    # closed forms l<=1
    assert oracle.spherical_harmonic(0, 0, 0.7, 0.2) == pytest.approx(0.5 / np.sqrt(np.pi))
    y11 = oracle.spherical_harmonic(1, 1, 0.7, 0.2)
    assert y11 == pytest.approx(-0.5 * np.sqrt(1.5 / np.pi) * np.sin(0.7) * np.exp(0.2j))


def test_amplitudes_and_compute_match_numpy(oracle):
    """scatter() + dsp() + store() + 1/NM (all_vectors_scatter_device.cpp:231-439) vs vectorised numpy"""
    NA, NF, NM = 50, 30, 11
    xyz = synth.trajectory(NF, NA, 20.0, 0.2, 5)
    b = synth.factors(NA)
    q = 1.4 * synth.unit_vectors(NM, 6)
    for dsp, method in (("autocorrelate", "fftw"), ("autocorrelate", "direct"), ("square", "fftw"), ("plain", "fftw")):
        fqt, fq, fq2, A = oracle.compute_all_vectors(xyz, b, q, dsp=dsp, method=method, return_amplitudes=True)
        Anp = oracle.np_amplitudes_all(xyz, b, q)
        assert rel_err(A, Anp) < 1e-13
        rfqt, rfq, rfq2 = oracle.np_dsp_store(Anp, dsp, method)
        assert rel_err(fqt, rfqt) < 1e-12 and abs(fq - rfq) < 1e-12 * abs(rfqt[0]) and abs(fq2 - rfq2) < 1e-12 * abs(rfq2)
    # threads (the reference's worker threads) and the frame-split ranks do not change the result
    base = oracle.compute_all_vectors(xyz, b, q)
    thr = oracle.compute_all_vectors(xyz, b, q, nthreads=4)
    fs = oracle.compute_all_vectors(xyz, b, q, nthreads=3, framesplit=True)
    assert np.array_equal(base[0], thr[0]) and np.array_equal(base[0], fs[0])


def test_self_matches_numpy(oracle):
    NA, NF, NM = 9, 25, 4
    xa = synth.trajectory(NF, NA, 20.0, 0.2, 7, layout=1)
    b = synth.factors(NA)
    q = 0.8 * synth.unit_vectors(NM, 8)
    fqt, fq, fq2 = oracle.compute_self_vectors(xa, b, q)
    # per-atom timelines: a[n,m,t] = b_n exp(i q_m . r_n(t))
    ph = np.einsum("ntc,mc->nmt", xa.astype(np.float64), q)
    T = (b[:, None, None] * np.exp(1j * ph)).reshape(NA * NM, NF)
    rfqt, rfq, rfq2 = oracle.np_dsp_store(T, norm=1.0 / NM)
    assert rel_err(fqt, rfqt) < 1e-12 and abs(fq - rfq) < 1e-12 * abs(rfqt[0]) and abs(fq2 - rfq2) < 1e-12 * abs(rfq2)


def test_mpsphere_matches_scipy(oracle):
    """multipole_scatter_device.cpp:467-497 with scipy's j_l and Y_lm; normalisation 1/(4 pi) (:395)"""
    NA, NF, L = 14, 5, 4
    xyz = synth.trajectory(NF, NA, 20.0, 0.3, 9, offset=-10.0)
    sph = oracle.cart_to_spherical(xyz)
    b = synth.factors(NA)
    mom = oracle.moments_sphere(L)
    ql = 0.9
    fqt, fq, fq2, A = oracle.compute_mpsphere(sph, b, ql, mom, dsp="square", return_amplitudes=True)
    r, phi, th = (sph[..., i].astype(np.float64) for i in range(3))
    Aref = np.array([[np.sum(4 * np.pi * (1j ** l) * b * sp.spherical_jn(l, ql * r[f]) * np.conj(sp.sph_harm_y(l, m, th[f], phi[f])))
                      for f in range(NF)] for l, m in mom])
    assert rel_err(A, Aref) < 1e-12
    rfqt, rfq, rfq2 = oracle.np_dsp_store(Aref, "square", norm=1 / (4 * np.pi))
    assert rel_err(fqt, rfqt) < 1e-12
    # cart -> spherical conversion (coor3d.cpp:168-215): theta = acos(z/r), phi in [0, 2 pi)
    x = xyz.astype(np.float64)
    assert np.allclose(r, np.linalg.norm(x, axis=-1), rtol=1e-6)
    assert np.all((phi >= 0) & (phi < 2 * np.pi + 1e-6)) and np.all((th >= 0) & (th <= np.pi))
    assert np.allclose(r * np.sin(th) * np.cos(phi), x[..., 0], atol=1e-4)
    with pytest.raises(RuntimeError):
        oracle.compute_mpsphere(sph, b, ql, [[1, 2]])


def test_known_answers(oracle):
    """KA1, KA2, KA3, KA5 of SURVEY 8c"""
    NF, bb = 40, 6.65
    static = np.tile(np.array([[1.5, -2.0, 0.25]], dtype=np.float32), (NF, 1, 1))
    fqt, fq, fq2 = oracle.compute_all_vectors(static, [bb], [[0.3, 0.1, -0.7]])
    assert np.allclose(fqt, bb ** 2, rtol=1e-13) and fq == pytest.approx(bb ** 2) and fq2 == pytest.approx(bb ** 4)
    # KA2: linear motion, fftw -> e^{+i q v tau}, direct -> e^{-i q v tau}
    v = 0.125
    lin = np.zeros((NF, 1, 3), dtype=np.float32)
    lin[:, 0, 0] = v * np.arange(NF)
    tau = np.arange(NF)
    f = oracle.compute_all_vectors(lin, [bb], [[0.8, 0, 0]], method="fftw")[0]
    d = oracle.compute_all_vectors(lin, [bb], [[0.8, 0, 0]], method="direct")[0]
    assert np.allclose(f, bb ** 2 * np.exp(1j * 0.8 * v * tau), atol=1e-12 * bb ** 2)
    assert np.allclose(d, bb ** 2 * np.exp(-1j * 0.8 * v * tau), atol=1e-12 * bb ** 2)
    # single atom, any trajectory: fq0 = b^2 exactly (up to rounding)
    walk = synth.trajectory(NF, 1, 10.0, 0.5, 3)
    assert oracle.compute_all_vectors(walk, [bb], [[0.4, 0.4, 0.1]])[0][0] == pytest.approx(bb ** 2, rel=1e-13)
    # KA5: square -> fq = mean_t |A|^2, fq0 = fqt[0]
    xyz = synth.trajectory(NF, 6, 10.0, 0.5, 4)
    b6 = synth.factors(6)
    q = [[0.4, 0.4, 0.1]]
    sq = oracle.compute_all_vectors(xyz, b6, q, dsp="square", return_amplitudes=True)
    assert sq[1] == pytest.approx(np.mean(np.abs(sq[3][0]) ** 2))
    # KA3: two static atoms, multipole sphere average -> Debye formula (to the float32 staging accuracy)
    d_, ql = 3.0, 1.1
    two = np.zeros((2, 2, 3), dtype=np.float32)
    two[:, 0] = (0.5, 0.25, -0.125)
    two[:, 1] = (0.5, 0.25 + d_, -0.125)
    bt = np.array([2.0, 3.5])
    exact = bt[0] ** 2 + bt[1] ** 2 + 2 * bt[0] * bt[1] * np.sin(ql * d_) / (ql * d_)
    mp = oracle.compute_mpsphere(oracle.cart_to_spherical(two), bt, ql, oracle.moments_sphere(14), dsp="square")
    assert mp[0][0].real == pytest.approx(exact, rel=1e-6)


def test_golden_fixtures(oracle):
    """the committed golden vectors still come out of the oracle bit-for-bit (to 1e-13)"""
    g = np.load(os.path.join(GOLD, "coherent_small.npz"))
    for dsp, method in (("autocorrelate", "fftw"), ("autocorrelate", "direct"), ("square", "fftw"), ("plain", "fftw")):
        for i, ql in enumerate(g["qls"]):
            fqt, fq, fq2 = oracle.compute_all_vectors(g["xyz"], g["b"], ql * g["u"], dsp=dsp, method=method)
            assert rel_err(fqt, g[f"all_{dsp}_{method}_{i}_fqt"]) < 1e-13
            assert np.allclose([fq, fq2], g[f"all_{dsp}_{method}_{i}_fq"], rtol=1e-12)
    g = np.load(os.path.join(GOLD, "self_small.npz"))
    for i, ql in enumerate(g["qls"]):
        fqt, fq, fq2 = oracle.compute_self_vectors(g["xyz_by_atom"], g["b"], ql * g["u"])
        assert rel_err(fqt, g[f"self_{i}_fqt"]) < 1e-13
    g = np.load(os.path.join(GOLD, "mpsphere_small.npz"))
    for i, ql in enumerate(g["qls"]):
        fqt, fq, fq2 = oracle.compute_mpsphere(oracle.cart_to_spherical(g["xyz"]), g["b"], ql, g["moments"])
        assert rel_err(fqt, g[f"mp_{i}_fqt"]) < 1e-13
    # the synthetic generator itself is part of the fixture contract
    assert np.array_equal(synth.trajectory(24, 40, 25.0, 0.3, 101), np.load(os.path.join(GOLD, "coherent_small.npz"))["xyz"])


def test_mpcylinder_c_port_matches_scipy_restatement(oracle):
    """MPCylinderScatterDevice::scatter (multipole_scatter_device.cpp:905-985): the C port (libm jn) against an independent
    numpy restatement with scipy.special.jv, several axes / q including the float-pi quadrant of the azimuth, moments up to
    l = 8 (Bessel orders 0..16); closed form for one atom on the axis."""
    from sassena_b200 import synth
    NA, NF = 40, 5
    xyz = synth.trajectory(NF, NA, 30.0, 0.2, 3, offset=-15.0)
    b = synth.factors(NA)
    mom = oracle.moments_cylinder(8)
    assert mom.shape == (33, 2) and mom[1].tolist() == [1, 0] and mom[-1].tolist() == [8, 3]
    for axis, q in [((0, 0, 1), (0.3, -0.4, 0.5)), ((1, 1, 0), (-0.7, 0.2, 0.1)), ((0, 1, 0), (0, 0, 0.9)),
                    ((0, 0, 1), (-0.5, -0.5, 0.0)), ((2, -1, 3), (1.5, 1.0, -2.0))]:
        cyl = oracle.cart_to_cylindrical(xyz, axis)
        assert cyl.dtype == np.float32 and np.all(cyl[..., 0] >= 0) and np.all((cyl[..., 1] >= 0) & (cyl[..., 1] <= 2 * np.pi))
        # the basis is orthonormal: r^2 + z^2 = |x|^2
        assert np.allclose(cyl[..., 0].astype(np.float64) ** 2 + cyl[..., 2].astype(np.float64) ** 2,
                           np.sum(xyz.astype(np.float64) ** 2, axis=-1), rtol=1e-5)
        *_, A = oracle.compute_mpcylinder(cyl, b, q, axis, mom, dsp="plain", return_amplitudes=True, nthreads=2)
        B = oracle.np_mpcylinder_amplitudes(cyl, b, q, axis, mom)
        assert np.max(np.abs(A - B)) < 1e-13 * np.max(np.abs(B))
    # one atom ON the axis: r = 0 -> only J_0 survives: A_(0,0) = sqrt(2 pi) b exp(i |z q_z|), all other moments vanish
    one = np.zeros((3, 1, 3), dtype=np.float32)
    one[:, 0, 2] = [0.0, 1.5, -2.5]
    *_, A = oracle.compute_mpcylinder(oracle.cart_to_cylindrical(one, (0, 0, 1)), [2.0], (0.3, 0.1, 0.7), (0, 0, 1), mom, dsp="plain",
                                      return_amplitudes=True)
    assert np.allclose(A[0], np.sqrt(2 * np.pi) * 2.0 * np.exp(1j * np.abs(one[:, 0, 2].astype(np.float64) * 0.7)), rtol=1e-14)
    assert np.max(np.abs(A[1:])) == 0.0
    with pytest.raises(RuntimeError):
        oracle.compute_mpcylinder(cyl, b, q, axis, [[0, 1]])
    with pytest.raises(RuntimeError):
        oracle.compute_mpcylinder(cyl, b, q, axis, [[2, 4]])


def test_orientational_averages_against_closed_forms(oracle):
    """KA7 of SURVEY 8c and its cylinder analogue: for a static cluster the orientational average of |A(q)|^2 has a closed
    form -- sphere: Debye, sum_ij b_i b_j sin(q r_ij)/(q r_ij); cylinder about z: sum_ij b_i b_j cos(q_z z_ij) J0(q_r rho_ij).
    The multipole devices must reproduce it to the float32 staging accuracy, the vector averages within their sampling error
    (random sphere vectors: Monte Carlo; equidistant cylinder angles: spectrally exact).  The atoms sit at z > 0 with q_z > 0,
    where the reference's exp(i |z q_z|) phase (multipole_scatter_device.cpp:941-947) coincides with exp(i z q_z)."""
    from scipy.special import j0
    rng = np.random.default_rng(3)
    NA = 6
    pos = (rng.normal(size=(NA, 3)) * 2.0).astype(np.float32)
    pos[:, 2] = np.abs(pos[:, 2]) + 0.5
    xyz = np.tile(pos, (2, 1, 1))
    b = np.array([2.0, -3.7, 6.6, 5.8, 1.0, 4.2])
    p64 = pos.astype(np.float64)
    # sphere
    ql = 1.2
    d = np.linalg.norm(p64[:, None, :] - p64[None, :, :], axis=-1)
    debye = float(np.sum(b[:, None] * b[None, :] * np.sinc(ql * d / np.pi)))
    mp = oracle.compute_mpsphere(oracle.cart_to_spherical(xyz), b, ql, oracle.moments_sphere(16), dsp="square")
    assert mp[0][0].real == pytest.approx(debye, rel=1e-6) and abs(mp[0][0].imag) < 1e-9
    assert mp[1].real == pytest.approx(debye, rel=1e-6)  # static: fq = fq0
    u = rng.normal(size=(20000, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    vec = oracle.compute_all_vectors(xyz, b, ql * u, dsp="square", nthreads=4)
    assert vec[0][0].real == pytest.approx(debye, rel=0.03)  # ~1/sqrt(20000) sampling error of |A|^2
    # cylinder about z
    q = np.array([0.9, 0.0, 0.7])
    axis = (0, 0, 1)
    rho = np.linalg.norm(p64[:, None, :2] - p64[None, :, :2], axis=-1)
    dz = p64[:, None, 2] - p64[None, :, 2]
    exact = float(np.sum(b[:, None] * b[None, :] * np.cos(q[2] * dz) * j0(0.9 * rho)))
    cyl = oracle.cart_to_cylindrical(xyz, axis)
    for L in (10, 20):
        mc = oracle.compute_mpcylinder(cyl, b, q, axis, oracle.moments_cylinder(L), dsp="square")
        assert mc[0][0].real == pytest.approx(exact, rel=1e-6)
    phi = np.linspace(0, 2 * np.pi, 720, endpoint=False)
    qv = np.stack([0.9 * np.cos(phi), 0.9 * np.sin(phi), np.full_like(phi, 0.7)], axis=1)
    vc = oracle.compute_all_vectors(xyz, b, qv, dsp="square", nthreads=4)
    assert vc[0][0].real == pytest.approx(exact, rel=1e-10)
    # the same vectors out of the reference's cylinder construction (abstract_vectors_scatter_device.cpp:130-149)
    sub = oracle.init_subvectors("cylinder", q, orient=np.stack([np.cos(phi), np.sin(phi), np.zeros_like(phi)], axis=1), axis=axis)
    vs = oracle.compute_all_vectors(xyz, b, sub, dsp="square", nthreads=4)
    assert vs[0][0].real == pytest.approx(exact, rel=1e-10)


def _ref_golden():
    return np.load(os.path.join(GOLD, "ref_smath.npz"))


def test_oracle_pinned_to_reference_build(oracle):
    """tests/golden/ref_smath.npz holds outputs of the REFERENCE's own src/math/smath.cpp, src/math/coor3d.cpp and
    src/decomposition/assignment.cpp, compiled where they lie (oracle/_ref/libsmath_ref.so, tests/golden/make_ref_golden.py).
    The oracle's restatements must reproduce them: the direct correlation, the square and the integer partitions bit for
    bit; the FFT form to rounding (the build's FFTW3 shim runs the oracle's DFT, so what is pinned there is the reference's
    padding / power spectrum / normalisation, not FFTW's arithmetic); the coordinate conversions after the stager's
    narrowing to float.  Also confirms SURVEY 8a-a4 on the reference's own code: its fftw_complex direct form is the
    complex conjugate of its FFT form, its std::vector overload is not."""
    g = _ref_golden()
    for i in range(7):
        x = g[f"corr_{i}_x"]
        d = oracle.auto_correlate_direct(x)
        assert np.array_equal(d, g[f"corr_{i}_direct"]), i
        f = oracle.auto_correlate_fftw(x)
        scale = np.max(np.abs(g[f"corr_{i}_fftw"]))
        assert np.max(np.abs(f - g[f"corr_{i}_fftw"])) <= 1e-15 * scale
        assert np.array_equal(x * np.conj(x), g[f"corr_{i}_square"]) or np.allclose(x * np.conj(x), g[f"corr_{i}_square"], rtol=1e-16)
        # a4: direct == conj(fftw) to rounding; the unused vector overload == fftw
        assert np.max(np.abs(g[f"corr_{i}_direct"] - np.conj(g[f"corr_{i}_fftw"]))) <= 1e-13 * scale
        assert np.max(np.abs(g[f"corr_{i}_direct_vec"] - g[f"corr_{i}_fftw"])) <= 1e-13 * scale
    pts = g["pts"].astype(np.float32)
    assert np.array_equal(oracle.cart_to_spherical(pts), g["sph"].astype(np.float32))
    for a, base, cyl in zip(g["axes"], g["bases"], g["cyl"]):
        assert np.array_equal(oracle.vector_base(a), base)
        assert np.array_equal(oracle.cart_to_cylindrical(pts, a), cyl.astype(np.float32))
    for mod, NN, rank, NAF, off, size, mx, first, last, total in g["assignments"]:
        got = (oracle.mod_assignment if mod else oracle.div_assignment)(int(NN), int(rank), int(NAF))
        assert got == (off, size, mx), (mod, NN, rank, NAF)
        if size:  # indices: Div = offset + i, Mod = rank + i * NN (assignment.cpp:36-50,91-105)
            idx = rank + NN * np.arange(size) if mod else off + np.arange(size)
            assert (idx[0], idx[-1], idx.sum()) == (first, last, total)


def test_host_layer_partitions_match_reference_build():
    """the PRODUCT's DivAssignment / ModAssignment (csrc/host/sassena_host.cpp) against the reference build's fixtures"""
    from sassena_b200 import host
    g = _ref_golden()
    for mod, NN, rank, NAF, off, size, mx, first, last, total in g["assignments"]:
        got = (host.mod_assignment if mod else host.div_assignment)(int(NN), int(rank), int(NAF))
        assert got == (off, size, mx), (mod, NN, rank, NAF)


def test_reference_build_live(oracle):
    """where oracle/_ref/libsmath_ref.so exists (built from /root/reference by `make -C oracle ref`): fresh random inputs
    through the reference's own code and through the oracle"""
    if not oracle.have_ref_smath():
        pytest.skip("oracle/_ref/libsmath_ref.so not built (no /root/reference on this machine)")
    rng = np.random.default_rng(7)
    for NF in (5, 64, 129, 1000):
        x = rng.normal(size=NF) + 1j * rng.normal(size=NF)
        assert np.array_equal(oracle.auto_correlate_direct(x), oracle.ref_auto_correlate_direct(x))
        r = oracle.ref_auto_correlate_fftw(x)
        assert np.max(np.abs(oracle.auto_correlate_fftw(x) - r)) <= 1e-15 * np.max(np.abs(r))
    pts = (rng.normal(size=(500, 3)) * 50).astype(np.float32)
    assert np.array_equal(oracle.cart_to_spherical(pts), oracle.ref_cart_to_spherical(pts.astype(np.float64)).astype(np.float32))
    for axis in ((0, 0, 1), (3, -2, 0.5), (1, 1, 1)):
        assert np.array_equal(oracle.cart_to_cylindrical(pts, axis),
                              oracle.ref_cart_to_cylindrical(pts.astype(np.float64), axis).astype(np.float32))
    for NN, NAF in ((5, 33), (8, 30000), (3, 2)):
        for rank in range(NN):
            assert oracle.div_assignment(NN, rank, NAF) == oracle.ref_assignment(False, NN, rank, NAF)[:3]
            assert oracle.mod_assignment(NN, rank, NAF) == oracle.ref_assignment(True, NN, rank, NAF)[:3]


def test_decomposition_pinned_to_reference_build(oracle):
    """penalties and plans recorded from the reference's own src/decomposition/decomposition_plan.cpp (oracle/_ref build):
    the oracle's restatement and the PRODUCT's DecompositionPlan (csrc/host/sassena_host.cpp) reproduce partitions, partition
    size and penalty for the automatic search and for manual partition sizes"""
    from sassena_b200 import host
    g = _ref_golden()
    for NN, NQ, NAF, NNpP, pen in g["penalties"]:
        assert oracle.decomposition_penalty(int(NN), int(NQ), int(NAF), int(NNpP)) == pen
    nauto = 0
    for NN, NQ, NAF, el, maxb, automatic, manual, part, psize, pen, colsum in g["plans"]:
        got = host.decomposition_plan(int(NN), int(NQ), int(NAF), int(el), int(maxb), 0.0, bool(automatic), int(manual))
        assert got == (part, psize, pen), (NN, NQ, NAF, maxb, automatic, manual)
        # colors = rank / partition size (decomposition_plan.cpp:163-184)
        assert int(((np.arange(NN) // psize) * (np.arange(NN) + 1)).sum()) == colsum
        if automatic:
            rc, opart, opsize, open_ = oracle.decomposition_plan(int(NN), int(NQ), int(NAF), int(el), int(maxb), 0.0)
            assert (rc, opart, opsize, open_) == (0, part, psize, pen)
            nauto += 1
    assert nauto > 100


def test_dcd_layout_pinned_to_reference_writer(tmp_path, oracle):
    """tests/golden/ref_writer.dcd was written by the reference's own DCDCoordinateWriter (src/stager/coordinate_writer.cpp,
    oracle/_ref build; two pieces, as two ranks of a partition write it).  The product's DCD writer reproduces the file
    byte for byte and its reader returns the coordinates that went in."""
    from sassena_b200 import host
    g = _ref_golden()
    xyz = g["dcd_xyz"]
    ref_bytes = open(os.path.join(GOLD, "ref_writer.dcd"), "rb").read()
    mine = str(tmp_path / "mine.dcd")
    host.write_dcd(mine, xyz)
    assert open(mine, "rb").read() == ref_bytes
    f = host.DCDFile(os.path.join(GOLD, "ref_writer.dcd"))
    assert (f.number_of_frames, f.number_of_atoms) == xyz.shape[:2]
    assert np.array_equal(f.read(), xyz)
    if oracle.have_ref_smath():  # live: other shapes
        rng = np.random.default_rng(11)
        for NF, NA in ((1, 1), (3, 50), (40, 7)):
            a = rng.normal(size=(NF, NA, 3)).astype(np.float32) * 20
            p = str(tmp_path / f"ref_{NF}_{NA}.dcd")
            oracle.ref_dcd_write(p, a)
            host.write_dcd(mine, a)
            assert open(mine, "rb").read() == open(p, "rb").read()
            assert np.array_equal(host.DCDFile(p).read(), a)
