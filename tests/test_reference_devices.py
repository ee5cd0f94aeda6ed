"""Parity against the REFERENCE's OWN scatter devices.  tests/golden/ref_devices.npz holds fqt / fq / fq2 written by the
reference's AllVectorsScatterDevice and SelfVectorsScatterDevice — its own amplitude loops, stagers, alignpad, DSP (smath.cpp),
store, final scaling, init_subvectors (file / sphere / cylinder / no averaging) — compiled where they lie for one MPI rank over
the shims in oracle/shim (tests/golden/make_ref_devices_golden.py).

* CPU: the oracle reproduces every case BIT FOR BIT (so the oracle IS the reference's arithmetic for these devices);
* GPU: the CUDA path, through the C-ABI, agrees with the reference's own output within the north-star tolerance 1e-9;
* where oracle/_ref exists, fresh inputs go through the reference devices live.

tests/golden/ref_multipole_devices.npz holds the same for the reference's MPSphereScatterDevice and MPCylinderScatterDevice
(multipole_scatter_device.cpp compiled where it lies; Boost.Math's three special functions served by the oracle's restatements,
which tests/test_oracle.py checks against scipy -- tests/golden/make_ref_multipole_golden.py)."""
import os

import numpy as np
import pytest

import sassena_b200
from sassena_b200 import synth

GOLD_MP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_multipole_devices.npz")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_devices.npz")
TOL = 1e-9


def _subvectors(oracle, g, vt, q):
    if vt == "none":
        return np.array([q])
    if vt == "cylinder":
        return oracle.init_subvectors("cylinder", q, orient=g["cyl"], axis=g["axis"])
    return oracle.init_subvectors("sphere", q, orient=g["u"])  # file and sphere: |q| * unit vector


def test_oracle_reproduces_reference_devices_bit_for_bit(oracle):
    g = np.load(GOLD)
    xyz, b = g["xyz"], g["b"]
    xa = np.ascontiguousarray(xyz.transpose(1, 0, 2))
    for k, (kind, vt, dsp, method) in enumerate(g["cases"]):
        for i, q in enumerate(g["qv"]):
            sub = _subvectors(oracle, g, vt, q)
            if kind == "all":
                fqt, fq, fq2 = oracle.compute_all_vectors(xyz, b, sub, dsp=dsp, method=method)
            else:
                fqt, fq, fq2 = oracle.compute_self_vectors(xa, b, sub, dsp=dsp, method=method)
            assert np.array_equal(fqt, g[f"case{k}_fqt"][i]), (kind, vt, dsp, method, i)
            assert fq == g[f"case{k}_fq"][i] and fq2 == g[f"case{k}_fq2"][i], (kind, vt, dsp, method, i)


def test_reference_devices_live(oracle):
    if not oracle.have_ref_smath():
        pytest.skip("oracle/_ref/libsmath_ref.so not built (no /root/reference on this machine)")
    NF, NA = 19, 31
    xyz = synth.trajectory(NF, NA, 18.0, 0.4, 77)
    b = synth.factors(NA)
    u = synth.unit_vectors(5, 9)
    qv = np.array([[1.1, -0.3, 0.2]])
    for kind in ("all", "self"):
        for threads in (1, 4):
            q, fqt, fq, fq2 = oracle.ref_scatter_run(kind, xyz, b, qv, orient=u, vectors_type="sphere", threads=threads)
            sub = np.linalg.norm(qv[0]) * u
            if kind == "all":
                r = oracle.compute_all_vectors(xyz, b, oracle.init_subvectors("sphere", qv[0], orient=u))
            else:
                r = oracle.compute_self_vectors(np.ascontiguousarray(xyz.transpose(1, 0, 2)), b,
                                                oracle.init_subvectors("sphere", qv[0], orient=u))
            assert np.array_equal(fqt[0], r[0]) and fq[0] == r[1] and fq2[0] == r[2], (kind, threads)
            assert np.allclose(sub, oracle.init_subvectors("sphere", qv[0], orient=u), rtol=1e-15)


def _qlen(q):  # CartesianCoor3D::length (coor3d.cpp): sqrt(x*x + y*y + z*z), left to right
    return float(np.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]))


def _oracle_multipole(oracle, g, kind, q, dsp, method, sph, cyl):
    if kind == "sphere":
        return oracle.compute_mpsphere(sph, g["b"], _qlen(q), g["mom_sphere"], dsp=dsp, method=method)
    return oracle.compute_mpcylinder(cyl, g["b"], q, g["axis"], g["mom_cylinder"], dsp=dsp, method=method)


def test_oracle_reproduces_reference_multipole_devices_bit_for_bit(oracle):
    """MPSphere / MPCylinder: moment loop, i^l and (-1)^l prefactors, conjugated Y_lm, atom summation order, sqrt(2 pi)
    normalisation, dsp, store and the 1/(4 pi) / 1/(2 pi) ... final scaling of the reference's own devices"""
    g = np.load(GOLD_MP)
    sph = oracle.cart_to_spherical(g["xyz"])
    cyl = oracle.cart_to_cylindrical(g["xyz"], g["axis"])
    for k, (kind, dsp, method) in enumerate(g["cases"]):
        for i, q in enumerate(g["qv"]):
            fqt, fq, fq2 = _oracle_multipole(oracle, g, kind, q, dsp, method, sph, cyl)
            assert np.array_equal(fqt, g[f"case{k}_fqt"][i]), (kind, dsp, method, i)
            assert fq == g[f"case{k}_fq"][i] and fq2 == g[f"case{k}_fq2"][i], (kind, dsp, method, i)


def test_reference_multipole_devices_live(oracle):
    if not oracle.have_ref_smath():
        pytest.skip("oracle/_ref/libsmath_ref.so not built (no /root/reference on this machine)")
    NF, NA = 14, 29
    xyz = synth.trajectory(NF, NA, 40.0, 0.6, 5, offset=-20.0)
    xyz[0, 0] = 0.0  # an atom at the origin
    xyz[1, 1, :2] = 0.0  # an atom on the z axis
    b = synth.factors(NA)
    qv = np.array([[0.2, -1.4, 0.5], [0.0, 0.0, 2.5]])
    axis = (0.0, 0.0, 1.0)
    # thread counts that divide the number of moments (144 and 33): the reference pads the last block of moments with the index
    # NM and its workers evaluate it (the `moment_index<NM` guard is commented out, multipole_scatter_device.cpp:199,671), i.e.
    # read multipole_index_[NM] out of bounds -- harmless garbage or a bare `throw;`, depending on the heap
    for threads in (1, 3):
        for kind, mom in (("sphere", oracle.moments_sphere(11)), ("cylinder", oracle.moments_cylinder(8))):
            q, fqt, fq, fq2 = oracle.ref_multipole_run(kind, xyz, b, qv, mom, axis=axis, threads=threads)
            for i in range(len(qv)):
                if kind == "sphere":
                    r = oracle.compute_mpsphere(oracle.cart_to_spherical(xyz), b, _qlen(qv[i]), mom)
                else:
                    r = oracle.compute_mpcylinder(oracle.cart_to_cylindrical(xyz, axis), b, qv[i], axis, mom)
                assert np.array_equal(fqt[i], r[0]) and fq[i] == r[1] and fq2[i] == r[2], (kind, threads, i)


@pytest.mark.gpu
def test_cuda_path_matches_reference_multipole_devices(oracle):
    """K5 / K6 through the C-ABI against the reference's own multipole devices' output, every committed case"""
    g = np.load(GOLD_MP)
    xyz, b, axis = g["xyz"], g["b"], g["axis"]
    worst = 0.0
    with sassena_b200.ScatterContext(0) as ctx:
        for k, (kind, dsp, method) in enumerate(g["cases"]):
            ctx.stage_frames(xyz)
            if kind == "sphere":
                ctx.frames_to_spherical()
            else:
                ctx.frames_to_cylindrical(axis)
            ctx.set_factors(b)
            for i, q in enumerate(g["qv"]):
                if kind == "sphere":
                    fqt, fq, fq2 = ctx.compute_mpsphere(_qlen(q), g["mom_sphere"], dsp=dsp, method=method)
                else:
                    fqt, fq, fq2 = ctx.compute_mpcylinder(q, axis, g["mom_cylinder"], dsp=dsp, method=method)
                rfqt, rfq, rfq2 = g[f"case{k}_fqt"][i], g[f"case{k}_fq"][i], g[f"case{k}_fq2"][i]
                scale = np.max(np.abs(rfqt))
                e = np.max(np.abs(fqt - rfqt)) / scale
                worst = max(worst, e)
                assert e < TOL, (kind, dsp, method, i, e)
                assert abs(fq - rfq) < TOL * scale and abs(fq2 - rfq2) <= TOL * max(abs(rfq2), 1e-300), (kind, dsp, method, i)
    print("worst fqt rel. err vs the reference's own multipole devices:", worst)


@pytest.mark.gpu
def test_cuda_path_matches_reference_devices(oracle):
    """the product (CUDA kernels through the C-ABI) against the reference's own devices' output, every committed case"""
    g = np.load(GOLD)
    xyz, b = g["xyz"], g["b"]
    xa = np.ascontiguousarray(xyz.transpose(1, 0, 2))
    worst = 0.0
    with sassena_b200.ScatterContext(0) as ctx:
        for k, (kind, vt, dsp, method) in enumerate(g["cases"]):
            if kind == "all":
                ctx.stage_frames(xyz)
            else:
                ctx.stage_atoms(xa)
            ctx.set_factors(b)
            for i, q in enumerate(g["qv"]):
                sub = _subvectors(oracle, g, vt, q)
                if kind == "all":
                    fqt, fq, fq2 = ctx.compute_all_vectors(sub, dsp=dsp, method=method)
                else:
                    fqt, fq, fq2 = ctx.compute_self_vectors(sub, dsp=dsp, method=method)
                rfqt, rfq, rfq2 = g[f"case{k}_fqt"][i], g[f"case{k}_fq"][i], g[f"case{k}_fq2"][i]
                scale = np.max(np.abs(rfqt))
                e = np.max(np.abs(fqt - rfqt)) / scale
                worst = max(worst, e)
                assert e < TOL, (kind, vt, dsp, method, i, e)
                assert abs(fq - rfq) < TOL * scale and abs(fq2 - rfq2) <= TOL * max(abs(rfq2), 1e-300), (kind, vt, dsp, method, i)
    print("worst fqt rel. err vs the reference's own devices:", worst)


@pytest.mark.parametrize("kind,nranks", [("all", 2), ("all", 3), ("self", 2), ("self", 3)])
def test_reference_multirank_branches_live(oracle, kind, nranks):
    """The reference's own NNPP > 1 code as the oracle: its devices on 2 and 3 ranks of one partition (threads over the
    shared-memory communicator of oracle/shim/boost/mpi.hpp) -- DivAssignment of the frames + all_to_all + alignpad
    (all_vectors_scatter_device.cpp:169-207,291-315), ModAssignment of the atoms + DataStagerByAtom's staged transposition
    (data_stager.cpp:249-338), the reductions to partition rank 0 -- give what its one-rank run and the oracle give (the sums
    over ranks come in another order, hence 1e-13 instead of bit for bit)."""
    if not oracle.have_ref_smath():
        pytest.skip("oracle/_ref/libsmath_ref.so not built (no /root/reference on this machine)")
    NF, NA = 23, 37  # neither divides by 2 or 3: ragged frame / atom blocks
    xyz = synth.trajectory(NF, NA, 20.0, 0.2, 3, offset=-10.0)
    b = synth.factors(NA)
    u = synth.unit_vectors(7, 1)
    qv = np.array([[0.7, 0.0, 0.0], [1.3, 0.2, 0.0]])
    for dsp, method in (("autocorrelate", "fftw"), ("autocorrelate", "direct"), ("square", "fftw")):
        one = oracle.ref_scatter_run(kind, xyz, b, qv, orient=u, dsp=dsp, method=method, threads=2)
        many = oracle.ref_scatter_run_ranks(kind, nranks, xyz, b, qv, orient=u, dsp=dsp, method=method, threads=2)
        assert np.array_equal(many[0], one[0])
        for i, q in enumerate(qv):
            sub = np.linalg.norm(q) * u
            if kind == "all":
                ofqt, ofq, ofq2 = oracle.compute_all_vectors(xyz, b, sub, dsp=dsp, method=method)
            else:
                ofqt, ofq, ofq2 = oracle.compute_self_vectors(np.ascontiguousarray(xyz.transpose(1, 0, 2)), b, sub, dsp=dsp, method=method)
            scale = np.max(np.abs(one[1][i]))
            assert np.max(np.abs(many[1][i] - one[1][i])) < 1e-13 * scale
            assert np.max(np.abs(many[1][i] - ofqt)) < 1e-13 * scale
            assert abs(many[2][i] - ofq) < 1e-13 * scale and abs(many[3][i] - ofq2) < 1e-13 * abs(ofq2)


@pytest.mark.gpu
def test_cuda_frame_sharded_path_matches_reference_multirank(oracle):
    """the product's frame-sharded coherent path (frame windows + amplitude exchange + DSP of a timeline block, here the three
    ranks' shares evaluated one after the other on one GPU and summed) against the REFERENCE's own three-rank run"""
    if not oracle.have_ref_smath():
        pytest.skip("oracle/_ref/libsmath_ref.so not built (no /root/reference on this machine)")
    import torch
    NF, NA, NM, NN = 47, 301, 13, 3
    xyz = synth.trajectory(NF, NA, 30.0, 0.2, 11)
    b = synth.factors(NA)
    u = synth.unit_vectors(NM, 4)
    qv = np.array([[0.9, 0.1, 0.0]])
    ref = oracle.ref_scatter_run_ranks("all", NN, xyz, b, qv, orient=u, threads=2)
    sub = np.linalg.norm(qv[0]) * u
    with sassena_b200.ScatterContext(0) as ctx:
        amp = torch.zeros(NM * NF * 2, dtype=torch.float64, device="cuda")
        tmp = torch.zeros_like(amp)
        for r in range(NN):  # every rank's frame block -> its columns of A (zero elsewhere); the sum assembles the timelines
            f0, f1 = (r * NF) // NN, ((r + 1) * NF) // NN
            ctx.stage_frames(np.ascontiguousarray(xyz[f0:f1]))
            ctx.set_frame_window(NF, f0)
            ctx.set_factors(b)
            ctx.all_vectors_amplitudes(sub, tmp.data_ptr())
            ctx.synchronize()
            amp += tmp
        torch.cuda.synchronize()
        plen = ctx.partial_len("autocorrelate")
        total = torch.zeros(plen, dtype=torch.float64, device="cuda")
        part = torch.zeros(plen, dtype=torch.float64, device="cuda")
        for r in range(NN):  # every rank's DivAssignment block of the timelines
            m0, m1 = (r * NM) // NN, ((r + 1) * NM) // NN
            ctx.all_vectors_dsp_partial(amp.data_ptr(), m0, m1 - m0, part.data_ptr())
            ctx.synchronize()
            total += part
        torch.cuda.synchronize()
        fqt, fq, fq2 = ctx.finalize(total.data_ptr(), 1.0 / NM)
    scale = np.max(np.abs(ref[1][0]))
    assert np.max(np.abs(fqt - ref[1][0])) < TOL * scale
    assert abs(fq - ref[2][0]) < TOL * scale and abs(fq2 - ref[3][0]) < TOL * abs(ref[3][0])
