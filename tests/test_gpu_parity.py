"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on identical inputs.

Tolerance: BASELINE.json north_star — 1e-9 relative on fq / fqt (checked norm-wise, see util.rel_err).
"""
import numpy as np
import pytest

import sassena_b200
from sassena_b200 import synth
from util import TOL, rel_err

pytestmark = pytest.mark.gpu


def small_case(NA=1000, NF=100, NM=100, seed=1, box=30.0, sigma=0.1):
    xyz = synth.trajectory(NF, NA, box, sigma, seed)
    b = synth.factors(NA)
    u = synth.unit_vectors(NM, seed + 1)
    return xyz, b, u


def test_synth_trajectory_twin(gpu_ctx):
    """device generator == numpy twin, bit for bit, in both layouts and for strided atom subsets"""
    NF, NA = 37, 301
    ref = synth.trajectory(NF, NA, 30.0, 0.1, 11)
    d = gpu_ctx.device_alloc(ref.nbytes)
    try:
        gpu_ctx.synth_trajectory(d, NF, NA, 30.0, 0.1, 11, layout=0)
        got = np.empty_like(ref)
        gpu_ctx.memcpy_d2h(got, d)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        gpu_ctx.synth_trajectory(d, NF, NA, 30.0, 0.1, 11, layout=1)
        got1 = np.empty((NA, NF, 3), dtype=np.float32)
        gpu_ctx.memcpy_d2h(got1, d)
        assert np.array_equal(got1, ref.transpose(1, 0, 2))
        # atoms 2, 7, 12, ... (ModAssignment style subset), offset box
        atoms = np.arange(2, NA, 5)
        ref2 = synth.trajectory(NF, NA, 30.0, 0.1, 11, atoms=atoms, offset=-15.0, layout=1)
        gpu_ctx.synth_trajectory(d, NF, NA, 30.0, 0.1, 11, layout=1, atom0=2, atom_stride=5, NA_out=len(atoms),
                                 offset=-15.0)
        got2 = np.empty_like(ref2)
        gpu_ctx.memcpy_d2h(got2, d)
        assert np.array_equal(got2, ref2)
    finally:
        gpu_ctx.device_free(d)


def test_amplitudes_config1(gpu_ctx, oracle):
    """K1 amplitudes A[m][f] vs all_vectors_scatter_device.cpp:418-438 on BASELINE config 1 (one |q|)."""
    xyz, b, u = small_case()
    q = 1.3 * u
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    gpu_ctx.compute_all_vectors(q)
    A = gpu_ctx.get_amplitudes(len(q))
    *_, Aref = oracle.compute_all_vectors(xyz, b, q, return_amplitudes=True, nthreads=4)
    assert rel_err(A, Aref) < 1e-12


@pytest.mark.parametrize("dsp,method", [("autocorrelate", "fftw"), ("autocorrelate", "direct"), ("square", "fftw"),
                                        ("plain", "fftw")])
def test_all_vectors_config1(gpu_ctx, oracle, dsp, method):
    """BASELINE config 1: 1k atoms, 100 frames, 10 |q| x 100 vectors, coherent fq/fq0/fq2/fqt."""
    xyz, b, u = small_case()
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    qls = synth.qlengths(0.2, 2.0, 10)
    nq = 10 if (dsp, method) == ("autocorrelate", "fftw") else 2
    for ql in qls[:nq]:
        q = oracle.init_subvectors("sphere", [ql, 0, 0], u)
        fqt, fq, fq2 = gpu_ctx.compute_all_vectors(q, dsp=dsp, method=method)
        rfqt, rfq, rfq2 = oracle.compute_all_vectors(xyz, b, q, dsp=dsp, method=method, nthreads=4)
        assert rel_err(fqt, rfqt) < TOL
        assert abs(fq - rfq) < TOL * abs(rfqt[0])
        assert abs(fq2 - rfq2) <= TOL * abs(rfq2)
        assert fqt[0] == pytest.approx(rfqt[0], rel=TOL)  # fq0
        # observed accuracy is far better than the contract
        assert rel_err(fqt, rfqt) < 1e-11


@pytest.mark.parametrize("NF", [1, 2, 3, 5, 64, 100, 129, 257, 1000])
def test_all_vectors_frame_counts(gpu_ctx, oracle, NF):
    """ragged / odd / power-of-two frame counts (zero padding 2NF -> L)"""
    NA, NM = 60, 7
    xyz = synth.trajectory(NF, NA, 20.0, 0.2, 3)
    b = synth.factors(NA)
    q = 0.9 * synth.unit_vectors(NM, 5)
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    fqt, fq, fq2 = gpu_ctx.compute_all_vectors(q)
    rfqt, rfq, rfq2 = oracle.compute_all_vectors(xyz, b, q)
    assert rel_err(fqt, rfqt) < TOL
    assert abs(fq - rfq) < TOL * abs(rfqt[0])
    assert abs(fq2 - rfq2) <= TOL * abs(rfq2)


@pytest.mark.parametrize("NA,NM", [(1, 1), (31, 8), (33, 9), (64, 64), (65, 65), (2000, 130)])
def test_all_vectors_ragged_atoms_vectors(gpu_ctx, oracle, NA, NM):
    """atom counts around the warp size, vector counts around the q-block (8) and CTA (64) sizes"""
    NF = 20
    xyz = synth.trajectory(NF, NA, 25.0, 0.3, 7)
    b = synth.factors(NA)
    q = 2.1 * synth.unit_vectors(NM, 8)
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    fqt, fq, fq2 = gpu_ctx.compute_all_vectors(q)
    A = gpu_ctx.get_amplitudes(NM)
    rfqt, rfq, rfq2, Aref = oracle.compute_all_vectors(xyz, b, q, return_amplitudes=True)
    assert rel_err(A, Aref) < 1e-12
    assert rel_err(fqt, rfqt) < TOL
    assert abs(fq - rfq) < TOL * abs(rfqt[0])


def test_large_phase_arguments(gpu_ctx, oracle):
    """|q.r| of several thousand radians: the range reduction must hold its accuracy"""
    NF, NA, NM = 8, 500, 16
    xyz = synth.trajectory(NF, NA, 2000.0, 1.0, 21, offset=-1000.0)
    b = synth.factors(NA)
    q = 5.0 * synth.unit_vectors(NM, 22)
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    gpu_ctx.compute_all_vectors(q, dsp="plain")
    A = gpu_ctx.get_amplitudes(NM)
    *_, Aref = oracle.compute_all_vectors(xyz, b, q, dsp="plain", return_amplitudes=True)
    # the reference's own rounding of p=x*qx+y*qy+z*qz (no FMA) is ~1e-12 rad here; stay well under 1e-9
    assert np.max(np.abs(A - Aref)) / (np.abs(b).sum()) < 1e-11


@pytest.mark.parametrize("dsp,method", [("autocorrelate", "fftw"), ("autocorrelate", "direct"), ("square", "fftw"),
                                        ("plain", "fftw")])
def test_self_vectors_small(gpu_ctx, oracle, dsp, method):
    """incoherent self scattering, per-atom timelines (self_vectors_scatter_device.cpp:145-239)"""
    NA, NF, NM = 40, 100, 12
    xyz = synth.trajectory(NF, NA, 30.0, 0.1, 3, layout=1)
    b = synth.factors(NA)
    q = 1.1 * synth.unit_vectors(NM, 4)
    gpu_ctx.stage_atoms(xyz)
    gpu_ctx.set_factors(b)
    fqt, fq, fq2 = gpu_ctx.compute_self_vectors(q, dsp=dsp, method=method)
    rfqt, rfq, rfq2 = oracle.compute_self_vectors(xyz, b, q, dsp=dsp, method=method, nthreads=4)
    assert rel_err(fqt, rfqt) < TOL
    assert abs(fq - rfq) < TOL * abs(rfqt[0])
    assert abs(fq2 - rfq2) <= TOL * abs(rfq2)


@pytest.mark.parametrize("NF", [1, 2, 3, 33, 100, 1000, 2049, 5000])
def test_self_vectors_frame_counts(gpu_ctx, oracle, NF):
    """fused self path: padded length L = R*N with N = 2^k <= 4096; NF up to several residues (R = 1, 2, 3)"""
    NA, NM = 3, 2
    xyz = synth.trajectory(NF, NA, 30.0, 0.1, 17, layout=1)
    b = synth.factors(NA)
    q = 1.3 * synth.unit_vectors(NM, 18)
    gpu_ctx.stage_atoms(xyz)
    gpu_ctx.set_factors(b)
    for method in ("fftw", "direct"):
        fqt, fq, fq2 = gpu_ctx.compute_self_vectors(q, method=method)
        rfqt, rfq, rfq2 = oracle.compute_self_vectors(xyz, b, q, method="fftw", nthreads=4)
        if method == "direct":
            rfqt, rfq = np.conj(rfqt), np.conj(rfq)
        assert rel_err(fqt, rfqt) < TOL
        assert abs(fq - rfq) < TOL * abs(rfqt[0])
        assert abs(fq2 - rfq2) <= TOL * abs(rfq2)


@pytest.mark.parametrize("NF,NA,NM", [(2049, 2, 3), (3000, 3, 2), (4096, 1, 5), (4097, 3, 2), (5000, 7, 5), (6200, 2, 3),
                                      (8193, 1, 1), (10000, 5, 40), (10000, 61, 7), (12289, 2, 3), (18000, 2, 2), (20000, 2, 3),
                                      (22000, 1, 2), (26000, 3, 1), (30000, 2, 2), (32768, 1, 2), (33000, 2, 2), (35000, 3, 3),
                                      (50000, 2, 3)])
def test_self_vectors_split_path(oracle, monkeypatch, NF, NA, NM):
    """2NF-1 > 2*4096 (R >= 3): the split path -- every frame evaluated once, R decimated sub-FFTs per timeline, an
    R-point DFT across them, power spectrum permuted back to the residue-major layout -- against the oracle, forced with
    SASSENA_SELF_PATH=split (the library picks it from R = 5 on and the fused kernel below; both are checked).
    NF = 10000 is BASELINE config 2's timeline length (R = 5), NF = 50000 config 5's (R = 25, two-stage 5 x 5 combine);
    33000 needs R = 17 and runs with R = 18 (3 x 6), 35000 R = 18; 18000 / 20000 / 30000 / 32768 are R = 9 (3 x 3), 10 (2 x 5),
    15 (3 x 5) and 16 (4 x 4), 22000 needs 11 and runs with 12 (3 x 4), 26000 needs 13 and runs with 14 (2 x 7); 2049 / 3000 / 4096 are R = 2 (sub-sequences of 1025 ... 2048
    frames: the shortest and the longest a 4096-point sub-transform takes).  Kernel A deals the (timeline, r) pairs out over
    296 CTAs: 1 x 1 x 5 = 5 pairs leave most CTAs idle, 61 x 7 x 5 = 2135 give every CTA 7 or 8 with shares that begin and end
    in the middle of a timeline."""
    xyz = synth.trajectory(NF, NA, 30.0, 0.1, 17, layout=1)
    b = synth.factors(NA)
    q = 1.3 * synth.unit_vectors(NM, 18)
    rfqt, rfq, rfq2 = oracle.compute_self_vectors(xyz, b, q, nthreads=8)
    for path in ("split", None):
        if path:
            monkeypatch.setenv("SASSENA_SELF_PATH", path)  # read when the plan for this NF is made
        else:
            monkeypatch.delenv("SASSENA_SELF_PATH", raising=False)
        with sassena_b200.ScatterContext(0) as ctx:
            ctx.stage_atoms(xyz)
            ctx.set_factors(b)
            fqt, fq, fq2 = ctx.compute_self_vectors(q)
        assert rel_err(fqt, rfqt) < TOL
        assert abs(fq - rfq) < TOL * abs(rfqt[0])
        assert abs(fq2 - rfq2) <= TOL * abs(rfq2)


@pytest.mark.parametrize("NF,NA,NM,W", [(9000, 9, 6, 4), (4500, 5, 3, 5), (50000, 3, 2, 2)])
def test_self_split_path_on_streamed_waves(oracle, NF, NA, NM, W):
    """the split path on wave buffers (sgpu_stage_atoms_prefetch / _swap): the coordinates stay in natural frame order there,
    so kernel A gathers its decimated sub-sequences with 4-byte cp.async instead of copying contiguous runs; partials
    accumulate over the waves (BASELINE config 5's path at timeline lengths that take R = 5, 3 and 25)."""
    xyz = synth.trajectory(NF, NA, 30.0, 0.1, 29, layout=1)
    b = synth.factors(NA)
    q = 1.1 * synth.unit_vectors(NM, 30)
    rfqt, rfq, rfq2 = oracle.compute_self_vectors(xyz, b, q, nthreads=8)
    with sassena_b200.ScatterContext(0) as ctx:
        host = ctx.pinned((NA, NF, 3), np.float32)
        host.array[:] = xyz
        plen = None
        ctx.stage_atoms_prefetch(host.array[0:min(W, NA)])
        for first in range(0, NA, W):
            cnt = min(W, NA - first)
            ctx.stage_atoms_swap()
            if first + cnt < NA:
                ctx.stage_atoms_prefetch(host.array[first + cnt:min(first + cnt + W, NA)])
            ctx.set_factors(b[first:first + cnt])
            if plen is None:
                plen = ctx.partial_len("autocorrelate")
                acc, part = ctx.device_alloc(plen * 8), ctx.device_alloc(plen * 8)
                ctx.compute_self_vectors_partial(q, acc)
            else:
                ctx.compute_self_vectors_partial(q, part)
                ctx.accumulate(acc, part, plen)
        fqt, fq, fq2 = ctx.finalize(acc, 1.0 / NM)
        ctx.synchronize()
        ctx.device_free(acc)
        ctx.device_free(part)
        host.free()
    assert rel_err(fqt, rfqt) < TOL
    assert abs(fq - rfq) < TOL * abs(rfqt[0])
    assert abs(fq2 - rfq2) <= TOL * abs(rfq2)


def test_self_split_layout_roundtrip_and_generic_combine(oracle):
    """the split path keeps owned atom-major coordinates in decimated frame order; switching to a dsp type that needs
    the natural order converts back, and again forward.  The generic (any R) combine kernel gives the same result."""
    import os
    NF, NA, NM = 9000, 5, 7
    xyz = synth.trajectory(NF, NA, 30.0, 0.1, 7, layout=1)
    b = synth.factors(NA)
    q = 0.8 * synth.unit_vectors(NM, 8)
    ctx = sassena_b200.ScatterContext(0)
    ctx.stage_atoms(xyz)
    ctx.set_factors(b)
    a1 = ctx.compute_self_vectors(q)
    s1 = ctx.compute_self_vectors(q, dsp="square")   # natural order again
    a2 = ctx.compute_self_vectors(q)                 # decimated again
    ctx.close()
    ra = oracle.compute_self_vectors(xyz, b, q, nthreads=8)
    rs = oracle.compute_self_vectors(xyz, b, q, dsp="square", nthreads=8)
    assert rel_err(a1[0], ra[0]) < TOL and rel_err(s1[0], rs[0]) < TOL and np.array_equal(a1[0], a2[0])
    os.environ["SASSENA_SELF_GENERIC_COMBINE"] = "1"
    try:
        ctx = sassena_b200.ScatterContext(0)
        ctx.stage_atoms(xyz)
        ctx.set_factors(b)
        g = ctx.compute_self_vectors(q)
        ctx.close()
    finally:
        del os.environ["SASSENA_SELF_GENERIC_COMBINE"]
    assert rel_err(g[0], a1[0]) < 1e-12 and abs(g[2] - a1[2]) <= 1e-12 * abs(a1[2])


def test_self_split_equals_fused_kernel(oracle):
    """the same job through both kernels (SASSENA_SELF_PATH selects at plan creation): packed partials agree to 1e-12"""
    import os
    NF, NA, NM = 9000, 6, 12
    xyz = synth.trajectory(NF, NA, 30.0, 0.1, 5, layout=1)
    b = synth.factors(NA)
    q = 0.8 * synth.unit_vectors(NM, 6)
    res = {}
    for path in ("fused", "split"):
        os.environ["SASSENA_SELF_PATH"] = path
        try:
            ctx = sassena_b200.ScatterContext(0)
            ctx.stage_atoms(xyz)
            ctx.set_factors(b)
            res[path] = ctx.compute_self_vectors(q)
            ctx.close()
        finally:
            del os.environ["SASSENA_SELF_PATH"]
    for a, c in zip(res["fused"], res["split"]):
        assert rel_err(np.atleast_1d(c), np.atleast_1d(a)) < 1e-12


def test_self_from_frames_modassignment(gpu_ctx, oracle):
    """frame-major input transposed on the GPU for the atoms of rank r under ModAssignment(NN, r, NA);
    the sum of the per-rank partials equals the single-rank result (self...:199-230 reduce)."""
    NA, NF, NM, NN = 23, 50, 5, 3
    frames = synth.trajectory(NF, NA, 30.0, 0.1, 9)
    b = synth.factors(NA)
    q = 0.7 * synth.unit_vectors(NM, 10)
    plen = None
    total = None
    for r in range(NN):
        gpu_ctx.stage_atoms_from_frames(frames, NN, r)
        off, size, _ = oracle.mod_assignment(NN, r, NA)
        assert gpu_ctx.NA == size
        ids = off + NN * np.arange(size)
        gpu_ctx.set_factors(b[ids])
        plen = gpu_ctx.partial_len("autocorrelate")
        d = gpu_ctx.device_alloc(plen * 8)
        gpu_ctx.compute_self_vectors_partial(q, d)
        gpu_ctx.synchronize()
        part = np.empty(plen)
        gpu_ctx.memcpy_d2h(part, d)
        gpu_ctx.device_free(d)
        total = part if total is None else total + part
    d = gpu_ctx.device_alloc(plen * 8)
    gpu_ctx.memcpy_h2d(d, total)
    fqt, fq, fq2 = gpu_ctx.finalize(d, 1.0 / NM)
    gpu_ctx.device_free(d)
    rfqt, rfq, rfq2 = oracle.compute_self_vectors(frames.transpose(1, 0, 2), b, q)
    assert rel_err(fqt, rfqt) < TOL
    assert abs(fq - rfq) < TOL * abs(rfqt[0])
    assert abs(fq2 - rfq2) <= TOL * abs(rfq2)


def test_all_vectors_sharded_equals_single(gpu_ctx, oracle):
    """KA6: q-vector sharding over k logical GPUs (DivAssignment of the NM subvectors) + summed partials
    == single-GPU result to 1e-12, == oracle to 1e-9."""
    xyz, b, u = small_case(NA=300, NF=64, NM=50)
    q = 1.7 * u
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    single = gpu_ctx.compute_all_vectors(q)
    for dsp in ("autocorrelate", "square"):
        single = gpu_ctx.compute_all_vectors(q, dsp=dsp)
        plen = gpu_ctx.partial_len(dsp)
        d = gpu_ctx.device_alloc(plen * 8)
        total = np.zeros(plen)
        for r in range(4):
            off, size, _ = oracle.div_assignment(4, r, len(q))
            gpu_ctx.compute_all_vectors_partial(q[off:off + size], d, dsp=dsp)
            gpu_ctx.synchronize()
            part = np.empty(plen)
            gpu_ctx.memcpy_d2h(part, d)
            total += part
        gpu_ctx.memcpy_h2d(d, total)
        fqt, fq, fq2 = gpu_ctx.finalize(d, 1.0 / len(q), dsp=dsp)
        gpu_ctx.device_free(d)
        assert rel_err(fqt, single[0]) < 1e-12
        assert abs(fq - single[1]) < 1e-12 * abs(single[0][0])
        rfqt, rfq, rfq2 = oracle.compute_all_vectors(xyz, b, q, dsp=dsp)
        assert rel_err(fqt, rfqt) < TOL


@pytest.mark.parametrize("nranks,NF", [(2, 64), (3, 50), (8, 41)])
def test_all_vectors_frame_sharded_equals_single(gpu_ctx, oracle, nranks, NF):
    """Frame decomposition (the reference's, all_vectors_scatter_device.cpp:61,170-205,248): k logical GPUs each stage
    a DivAssignment block of the frames, fill their columns of A[NM][NF], the buffers are summed (the all-reduce),
    each rank correlates a block of the timelines, the partials are summed and finalized.  == single-GPU result to
    1e-12 and == oracle to 1e-9, for every dsp type; ragged blocks (NF not divisible) included."""
    xyz, b, u = small_case(NA=203, NF=NF, NM=37)
    q = 1.3 * u
    NM = len(q)
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    singles = {dsp: gpu_ctx.compute_all_vectors(q, dsp=dsp) for dsp in ("autocorrelate", "square", "plain")}
    d_amp = gpu_ctx.device_alloc(NM * NF * 16)
    A = np.zeros((NM, NF), dtype=np.complex128)
    part = np.empty((NM, NF), dtype=np.complex128)
    for r in range(nranks):
        off, size, _ = oracle.div_assignment(nranks, r, NF)
        gpu_ctx.stage_frames(xyz[off:off + size])
        gpu_ctx.set_frame_window(NF, off)
        gpu_ctx.set_factors(b)
        gpu_ctx.all_vectors_amplitudes(q, d_amp)
        gpu_ctx.synchronize()
        gpu_ctx.memcpy_d2h(part.view(np.float64), d_amp)
        assert np.all(part[:, :off] == 0) and np.all(part[:, off + size:] == 0)
        A += part
    # the fused single-GPU call refuses a windowed state instead of silently correlating a fragment
    with pytest.raises(sassena_b200.SgpuError):
        gpu_ctx.compute_all_vectors(q)
    ref_amp = oracle.compute_all_vectors(xyz, b, q, dsp="plain", return_amplitudes=True)[-1]
    assert np.max(np.abs(A - ref_amp)) < 1e-11 * np.max(np.abs(ref_amp))
    gpu_ctx.memcpy_h2d(d_amp, A.view(np.float64))
    for dsp, single in singles.items():
        plen = gpu_ctx.partial_len(dsp)
        d = gpu_ctx.device_alloc(plen * 8)
        total = np.zeros(plen)
        for r in range(nranks):
            off, size, _ = oracle.div_assignment(nranks, r, NM)
            gpu_ctx.all_vectors_dsp_partial(d_amp, off, size, d, dsp=dsp)
            gpu_ctx.synchronize()
            p = np.empty(plen)
            gpu_ctx.memcpy_d2h(p, d)
            total += p
        gpu_ctx.memcpy_h2d(d, total)
        fqt, fq, fq2 = gpu_ctx.finalize(d, 1.0 / NM, dsp=dsp)
        gpu_ctx.device_free(d)
        assert len(fqt) == NF
        assert rel_err(fqt, single[0]) < 1e-12
        assert abs(fq - single[1]) < 1e-12 * abs(single[0][0]) and abs(fq2 - single[2]) < 1e-12 * abs(single[2])
        rfqt, rfq, rfq2 = oracle.compute_all_vectors(xyz, b, q, dsp=dsp)
        assert rel_err(fqt, rfqt) < TOL
    gpu_ctx.device_free(d_amp)


@pytest.mark.parametrize("NQ", [1, 3, 4, 8, 16, 21, 37, 61])
def test_all_vectors_scan_matches_per_q(gpu_ctx, oracle, NQ):
    """|q|-scan kernel (two sincos + a 3-term recurrence per (atom, direction)): every |q| of an exactly spaced batch
    equals the per-|q| GPU result to 1e-11 and the oracle to 1e-9; NQ values exercise every pass size and masking."""
    xyz, b, u = small_case(NA=301, NF=40, NM=37)
    s = 0.3 + 0.17 * np.arange(NQ)
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    fqt, fq, fq2 = gpu_ctx.compute_all_vectors_scan(u, s)
    plain, corr, single = gpu_ctx.last_scan_plan()
    assert corr == 0 and ((plain >= 1 and single == 0) if NQ >= 3 else single == NQ)
    assert fqt.shape == (NQ, 40)
    for n in sorted({0, NQ // 2, NQ - 1}):
        q = s[n] * u
        g = gpu_ctx.compute_all_vectors(q)
        assert rel_err(fqt[n], g[0]) < 1e-11
        assert abs(fq[n] - g[1]) < 1e-11 * abs(g[0][0]) and abs(fq2[n] - g[2]) < 1e-11 * abs(g[2])
        r = oracle.compute_all_vectors(xyz, b, q)
        assert rel_err(fqt[n], r[0]) < TOL
        assert abs(fq[n] - r[1]) < TOL * abs(r[0][0]) and abs(fq2[n] - r[2]) < TOL * abs(r[2])


def test_all_vectors_scan_float_rounded_scan(gpu_ctx, oracle):
    """The reference builds scans from float-rounded fractions (parameters.cpp:1151): the |q| are equally spaced only to
    ~1e-8.  Such batches take the corrected kernel (first order in FP64, second order in FP32) and still equal the
    oracle evaluated at the exact q-vectors; a large box makes the phase deviations as big as they get in practice."""
    from sassena_b200 import host
    xyz, b, u = small_case(NA=257, NF=24, NM=19, box=400.0)
    qv = host.create_from_scans([{"base": (1, 0, 0), "from": 0.1, "to": 5.0, "points": 50}])
    s = np.linalg.norm(qv, axis=1)
    assert np.max(np.abs(np.diff(s, 2))) > 1e-9  # not a progression
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    fqt, fq, fq2 = gpu_ctx.compute_all_vectors_scan(u, s)
    plain, corr, single = gpu_ctx.last_scan_plan()
    assert plain == 0 and corr >= 3 and single == 0
    worst = 0.0
    for n in range(0, 50, 7):
        r = oracle.compute_all_vectors(xyz, b, s[n] * u)
        worst = max(worst, rel_err(fqt[n], r[0]))
        assert abs(fq[n] - r[1]) < TOL * abs(r[0][0])
    assert worst < 1e-10, worst  # an order of magnitude inside the tolerance
    # uncorrected, the same batch would miss the tolerance: the deviation is real
    snapped = np.linspace(s[0], s[-1], 50)
    r = oracle.compute_all_vectors(xyz, b, snapped[24] * u)
    assert rel_err(fqt[24], r[0]) > 1e-8


def test_all_vectors_scan_float_rounded_scan_fp32_first_order(oracle, monkeypatch):
    """A float-rounded 50-point scan whose phase deviations are small (box 60: theta_max ~ 1.6e-5) qualifies for the corrected
    kernel that sums the first-order term in packed FP32 as well (scan_sym.cu, CORR = 2: 25-|q| passes, 3 FP64 instructions
    per evaluation).  plan_scan admits it only while its error bound theta_max (K^2 ulp32 + accumulation) is below 5e-10;
    measured here against the oracle at the exact q-vectors: ~1e-11, tolerance 1e-9.  SASSENA_SCAN_FP64_CORR keeps the
    first-order sums in FP64 (three passes)."""
    from sassena_b200 import host
    xyz, b, u = small_case(NA=516, NF=12, NM=19, box=60.0)
    qv = host.create_from_scans([{"base": (1, 0, 0), "from": 0.1, "to": 5.0, "points": 50}])
    s = np.linalg.norm(qv, axis=1)
    res = {}
    for fp64 in (False, True):
        if fp64:
            monkeypatch.setenv("SASSENA_SCAN_FP64_CORR", "1")
        with sassena_b200.ScatterContext(0) as ctx:
            ctx.stage_frames(xyz)
            ctx.set_factors(b)
            res[fp64] = ctx.compute_all_vectors_scan(u, s)
            plain, corr, single = ctx.last_scan_plan()
        assert plain == 0 and single == 0 and corr == (3 if fp64 else 2)
    worst = 0.0
    for n in range(50):
        r = oracle.compute_all_vectors(xyz, b, s[n] * u)
        worst = max(worst, rel_err(res[False][0][n], r[0]))
        assert rel_err(res[False][0][n], r[0]) < TOL and rel_err(res[True][0][n], r[0]) < 1e-11
        assert abs(res[False][1][n] - r[1]) < TOL * abs(r[0][0])
    assert worst < 2e-10, worst  # the bound plan_scan works with is 5e-10; typical 1e-11


def test_all_vectors_scan_arbitrary_spacing_falls_back(gpu_ctx, oracle):
    """|q| values with no progression at all: the call is still valid, each |q| runs through the general kernel"""
    xyz, b, u = small_case(NA=120, NF=16, NM=12)
    s = np.array([0.2, 0.25, 0.7, 0.71, 1.9, 3.3])
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    fqt, fq, fq2 = gpu_ctx.compute_all_vectors_scan(u, s)
    assert gpu_ctx.last_scan_plan() == (0, 0, len(s))
    for n in range(len(s)):
        r = oracle.compute_all_vectors(xyz, b, s[n] * u)
        assert rel_err(fqt[n], r[0]) < TOL


@pytest.mark.parametrize("dsp", ["square", "plain"])
def test_all_vectors_scan_other_dsp_and_unaligned_atoms(gpu_ctx, oracle, dsp):
    """NA % 4 != 0 takes the cp.async staging path; square / plain dsp; long recurrence far from the origin"""
    xyz, b, u = small_case(NA=203, NF=24, NM=19, box=300.0)
    s = 1.1 + 0.45 * np.arange(16)
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    fqt, fq, fq2 = gpu_ctx.compute_all_vectors_scan(u, s, dsp=dsp)
    for n in (0, 7, 15):
        r = oracle.compute_all_vectors(xyz, b, s[n] * u, dsp=dsp)
        assert rel_err(fqt[n], r[0]) < TOL
        assert abs(fq[n] - r[1]) < TOL * max(abs(r[1]), abs(r[0]).max())


@pytest.mark.parametrize("NA", [160, 203])
def test_all_vectors_scan_per_q_factors(gpu_ctx, oracle, NA):
    """|q|-dependent factors (X-ray form factors / background): equally spaced |q| take the scan kernel with one factor row
    per |q| staged next to the coordinates (NA % 4 == 0: TMA rows, otherwise cp.async); a float-rounded scan with
    differing rows falls back to the general kernel per |q|; identical rows count as uniform.  All equal the oracle."""
    xyz, b, u = small_case(NA=NA, NF=16, NM=23)
    for NQ in (5, 27):
        s = 0.2 + 0.3 * np.arange(NQ)
        gpu_ctx.stage_frames(xyz)
        bq = np.stack([b * (1.0 + 0.1 * n) - 0.05 * n for n in range(NQ)])
        gpu_ctx.set_factors_batch(bq)
        fqt, fq, fq2 = gpu_ctx.compute_all_vectors_scan(u, s)
        plain, corr, single = gpu_ctx.last_scan_plan()
        assert plain >= 1 and corr == 0 and single == 0
        for n in range(NQ):
            r = oracle.compute_all_vectors(xyz, bq[n], s[n] * u)
            assert rel_err(fqt[n], r[0]) < TOL
            assert abs(fq[n] - r[1]) < TOL * abs(r[0][0])
    NQ = 5
    s = 0.2 + 0.3 * np.arange(NQ)
    bq = bq[:NQ]
    gpu_ctx.set_factors_batch(bq)
    s_rounded = s * (1 + 1e-8 * np.array([0, 1, -1, 2, 0]))
    fqt, fq, fq2 = gpu_ctx.compute_all_vectors_scan(u, s_rounded)
    assert gpu_ctx.last_scan_plan() == (0, 0, NQ)
    for n in range(NQ):
        r = oracle.compute_all_vectors(xyz, bq[n], s_rounded[n] * u)
        assert rel_err(fqt[n], r[0]) < TOL
    gpu_ctx.set_factors_batch(np.stack([b] * NQ))
    fqt, fq, fq2 = gpu_ctx.compute_all_vectors_scan(u, s)
    assert gpu_ctx.last_scan_plan() == (1, 0, 0)
    for n in range(NQ):
        r = oracle.compute_all_vectors(xyz, b, s[n] * u)
        assert rel_err(fqt[n], r[0]) < TOL


def test_all_vectors_scan_frame_window(gpu_ctx, oracle):
    """scan amplitudes of a frame block land in the right columns of A[NQ][NM][NF_total]"""
    xyz, b, u = small_case(NA=128, NF=30, NM=11)
    NQ, NM, NF = 6, 11, 30
    s = 0.4 + 0.25 * np.arange(NQ)
    d_amp = gpu_ctx.device_alloc(NQ * NM * NF * 16)
    A = np.zeros((NQ, NM, NF), dtype=np.complex128)
    part = np.empty_like(A)
    for r in range(3):
        off, size, _ = oracle.div_assignment(3, r, NF)
        gpu_ctx.stage_frames(xyz[off:off + size])
        gpu_ctx.set_frame_window(NF, off)
        gpu_ctx.set_factors(b)
        gpu_ctx.all_vectors_scan_amplitudes(u, s, d_amp)
        gpu_ctx.synchronize()
        gpu_ctx.memcpy_d2h(part.view(np.float64), d_amp)
        assert np.all(part[:, :, :off] == 0) and np.all(part[:, :, off + size:] == 0)
        A += part
    gpu_ctx.device_free(d_amp)
    for n in range(NQ):
        ref = oracle.compute_all_vectors(xyz, b, s[n] * u, dsp="plain", return_amplitudes=True)[-1]
        assert np.max(np.abs(A[n] - ref)) < 1e-11 * np.max(np.abs(ref))


@pytest.mark.parametrize("dsp", ["autocorrelate", "square"])
def test_mpsphere_small(gpu_ctx, oracle, dsp):
    """multipole sphere moments (multipole_scatter_device.cpp:428-498), L=6, cartesian input converted on the GPU"""
    NA, NF, L = 300, 16, 6
    xyz = synth.trajectory(NF, NA, 40.0, 0.3, 13, offset=-20.0)
    b = synth.factors(NA)
    mom = oracle.moments_sphere(L)
    sph = oracle.cart_to_spherical(xyz)
    for ql in (0.05, 0.4, 1.5):
        gpu_ctx.stage_frames(xyz)
        gpu_ctx.frames_to_spherical()
        gpu_ctx.set_factors(b)
        fqt, fq, fq2 = gpu_ctx.compute_mpsphere(ql, mom, dsp=dsp)
        A = gpu_ctx.get_amplitudes(len(mom))
        rfqt, rfq, rfq2, Aref = oracle.compute_mpsphere(sph, b, ql, mom, dsp=dsp, nthreads=4, return_amplitudes=True)
        assert rel_err(A, Aref) < 1e-9
        assert rel_err(fqt, rfqt) < TOL
        assert abs(fq - rfq) < TOL * abs(rfqt[0])
        assert abs(fq2 - rfq2) <= TOL * abs(rfq2)


def test_mpsphere_resolution20_staged_spherical(gpu_ctx, oracle):
    """default multipole.moments.resolution=20 (441 moments), input staged already in (r,phi,theta)"""
    NA, NF = 150, 6
    xyz = synth.trajectory(NF, NA, 60.0, 0.3, 15, offset=-30.0)
    b = synth.factors(NA)
    mom = oracle.moments_sphere(20)
    sph = oracle.cart_to_spherical(xyz)
    from sassena_b200 import REPR_SPHERICAL
    gpu_ctx.stage_frames(sph, repr=REPR_SPHERICAL)
    gpu_ctx.set_factors(b)
    fqt, fq, fq2 = gpu_ctx.compute_mpsphere(0.3, mom, dsp="square")
    rfqt, rfq, rfq2 = oracle.compute_mpsphere(sph, b, 0.3, mom, dsp="square", nthreads=8)
    assert rel_err(fqt, rfqt) < TOL
    assert abs(fq - rfq) < TOL * abs(rfqt[0])


def test_known_answers(gpu_ctx):
    """KA1/KA2 of SURVEY 8c: single static atom -> fqt = b^2, fq = b^2, fq2 = b^4; single atom in linear
    motion, no orientational average -> fftw: b^2 e^{+i q.v tau}, direct: conjugate."""
    NF = 50
    b = np.array([6.65])
    xyz = np.tile(np.array([[1.5, -2.0, 0.25]], dtype=np.float32), (NF, 1, 1))
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    fqt, fq, fq2 = gpu_ctx.compute_all_vectors(np.array([[0.3, 0.1, -0.7]]))
    assert np.allclose(fqt, b[0] ** 2, rtol=1e-12, atol=0)
    assert fq == pytest.approx(b[0] ** 2, rel=1e-12)
    assert fq2 == pytest.approx(b[0] ** 4, rel=1e-12)
    # linear motion along x with exactly representable steps
    v = 0.125
    xyz = np.zeros((NF, 1, 3), dtype=np.float32)
    xyz[:, 0, 0] = v * np.arange(NF)
    q = np.array([[0.8, 0.0, 0.0]])
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    tau = np.arange(NF)
    fqt, _, _ = gpu_ctx.compute_all_vectors(q, method="fftw")
    assert np.allclose(fqt, b[0] ** 2 * np.exp(1j * 0.8 * v * tau), rtol=0, atol=1e-11 * b[0] ** 2)
    fqt, _, _ = gpu_ctx.compute_all_vectors(q, method="direct")
    assert np.allclose(fqt, b[0] ** 2 * np.exp(-1j * 0.8 * v * tau), rtol=0, atol=1e-11 * b[0] ** 2)


def test_cluster_averages_closed_forms(gpu_ctx, oracle):
    """KA7 on the device: for a static cluster the multipole sphere device gives the Debye sum, the multipole cylinder
    device sum_ij b_i b_j cos(q_z z_ij) J0(q_r rho_ij) (atoms at z > 0, q_z > 0: there the reference's exp(i|z q_z|) phase
    equals exp(i z q_z)), to the float32 staging accuracy; equidistant cylinder vectors reproduce the latter exactly."""
    from scipy.special import j0
    rng = np.random.default_rng(3)
    NA = 6
    pos = (rng.normal(size=(NA, 3)) * 2.0).astype(np.float32)
    pos[:, 2] = np.abs(pos[:, 2]) + 0.5
    xyz = np.tile(pos, (3, 1, 1))
    b = np.array([2.0, -3.7, 6.6, 5.8, 1.0, 4.2])
    p64 = pos.astype(np.float64)
    ql = 1.2
    d = np.linalg.norm(p64[:, None, :] - p64[None, :, :], axis=-1)
    debye = float(np.sum(b[:, None] * b[None, :] * np.sinc(ql * d / np.pi)))
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.frames_to_spherical()
    gpu_ctx.set_factors(b)
    fqt, fq, _ = gpu_ctx.compute_mpsphere(ql, oracle.moments_sphere(16), dsp="square")
    assert fqt[0].real == pytest.approx(debye, rel=1e-6) and fq.real == pytest.approx(debye, rel=1e-6)
    q = np.array([0.9, 0.0, 0.7])
    axis = (0, 0, 1)
    rho = np.linalg.norm(p64[:, None, :2] - p64[None, :, :2], axis=-1)
    dz = p64[:, None, 2] - p64[None, :, 2]
    exact = float(np.sum(b[:, None] * b[None, :] * np.cos(q[2] * dz) * j0(0.9 * rho)))
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.frames_to_cylindrical(axis)
    gpu_ctx.set_factors(b)
    fqt, _, _ = gpu_ctx.compute_mpcylinder(q, axis, oracle.moments_cylinder(12), dsp="square")
    assert fqt[0].real == pytest.approx(exact, rel=1e-6)
    phi = np.linspace(0, 2 * np.pi, 720, endpoint=False)
    qv = np.stack([0.9 * np.cos(phi), 0.9 * np.sin(phi), np.full_like(phi, 0.7)], axis=1)
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    fqt, _, _ = gpu_ctx.compute_all_vectors(qv, dsp="square")
    assert fqt[0].real == pytest.approx(exact, rel=1e-10)


def test_two_atom_debye(gpu_ctx, oracle):
    """KA3: two static atoms at distance d; sphere average -> b1^2+b2^2+2 b1 b2 sin(qd)/(qd).
    The multipole expansion (L=14) reproduces it to the accuracy the reference's float32 (r,phi,theta) staging
    allows (~1e-7, data_stager.cpp:111-113); 4000 random vectors agree within Monte-Carlo error."""
    d, ql = 3.0, 1.1
    b = np.array([2.0, 3.5])
    xyz = np.zeros((4, 2, 3), dtype=np.float32)
    xyz[:, 0] = (0.5, 0.25, -0.125)
    xyz[:, 1] = (0.5, 0.25 + d, -0.125)
    exact = b[0] ** 2 + b[1] ** 2 + 2 * b[0] * b[1] * np.sin(ql * d) / (ql * d)
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.frames_to_spherical()
    gpu_ctx.set_factors(b)
    fqt, fq, fq2 = gpu_ctx.compute_mpsphere(ql, oracle.moments_sphere(14), dsp="square")
    assert fqt[0].real == pytest.approx(exact, rel=1e-6)
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    fqt, _, _ = gpu_ctx.compute_all_vectors(ql * synth.unit_vectors(4000, 3), dsp="square")
    assert fqt[0].real == pytest.approx(exact, rel=0.05)


def test_error_paths(gpu_ctx):
    """C-ABI error behaviour mirrors the reference's Err::write+throw sites, as return codes."""
    import sassena_b200
    ctx = sassena_b200.ScatterContext(0)
    with pytest.raises(sassena_b200.SgpuError) as e:
        ctx.compute_all_vectors(np.ones((1, 3)))  # nothing staged
    assert e.value.code == 3
    xyz, b, u = small_case(NA=10, NF=4, NM=2)
    ctx.stage_frames(xyz)
    with pytest.raises(sassena_b200.SgpuError):
        ctx.compute_all_vectors(u)  # factors missing
    ctx.set_factors(b)
    with pytest.raises(sassena_b200.SgpuError) as e:
        ctx.compute_all_vectors(u, dsp=7)  # "DSP type not understood"
    assert e.value.code == 1 and "DSP type not understood" in e.value.msg
    with pytest.raises(sassena_b200.SgpuError) as e:
        ctx.compute_all_vectors(u, method=5)
    assert "Correlation method not understood" in e.value.msg
    with pytest.raises(sassena_b200.SgpuError) as e:
        ctx.compute_mpsphere(1.0, [[1, 2]])  # frames not spherical
    ctx.frames_to_spherical()
    with pytest.raises(sassena_b200.SgpuError) as e:
        ctx.compute_mpsphere(1.0, [[1, 2]])  # |m| > l
    assert "Combination of Major and minor moment not allowed" in e.value.msg
    with pytest.raises(sassena_b200.SgpuError):
        ctx.compute_self_vectors(u)  # atoms not staged
    ctx.close()


def test_launch_counter_and_timers(gpu_ctx):
    xyz, b, u = small_case(NA=100, NF=16, NM=8)
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    n0 = gpu_ctx.launch_count
    gpu_ctx.compute_all_vectors(u)
    assert gpu_ctx.launch_count > n0
    assert gpu_ctx.last_amplitude_ms() > 0
    assert gpu_ctx.last_dsp_ms() > 0


def test_golden_fixtures_gpu(gpu_ctx):
    """the CUDA path against the committed golden vectors (tests/golden, generated by the pinned oracle)"""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gold, "coherent_small.npz"))
    gpu_ctx.stage_frames(g["xyz"])
    gpu_ctx.set_factors(g["b"])
    for dsp, method in (("autocorrelate", "fftw"), ("autocorrelate", "direct"), ("square", "fftw"), ("plain", "fftw")):
        for i, ql in enumerate(g["qls"]):
            fqt, fq, fq2 = gpu_ctx.compute_all_vectors(ql * g["u"], dsp=dsp, method=method)
            assert rel_err(fqt, g[f"all_{dsp}_{method}_{i}_fqt"]) < TOL
            ref = g[f"all_{dsp}_{method}_{i}_fq"]
            assert abs(fq - ref[0]) < TOL * abs(g[f"all_{dsp}_{method}_{i}_fqt"][0]) and abs(fq2 - ref[1]) <= TOL * abs(ref[1])
    g = np.load(os.path.join(gold, "self_small.npz"))
    gpu_ctx.stage_atoms(g["xyz_by_atom"])
    gpu_ctx.set_factors(g["b"])
    for i, ql in enumerate(g["qls"]):
        fqt, fq, fq2 = gpu_ctx.compute_self_vectors(ql * g["u"])
        assert rel_err(fqt, g[f"self_{i}_fqt"]) < TOL
    g = np.load(os.path.join(gold, "mpsphere_small.npz"))
    gpu_ctx.stage_frames(g["xyz"])
    gpu_ctx.frames_to_spherical()
    gpu_ctx.set_factors(g["b"])
    for i, ql in enumerate(g["qls"]):
        fqt, fq, fq2 = gpu_ctx.compute_mpsphere(ql, g["moments"])
        assert rel_err(fqt, g[f"mp_{i}_fqt"]) < TOL


def test_host_layer_devices_on_gpu(oracle):
    """ScatterDeviceFactory::create + device.run() (C++ host layer) over the real CUDA backend: all three devices,
    several |q| through the runner loop, results through the IResultSink callback."""
    from sassena_b200 import host
    NA, NF = 200, 48
    xyz = synth.trajectory(NF, NA, 30.0, 0.2, 31, offset=-15.0)
    b = synth.factors(NA)
    qv = host.create_from_scans([{"base": (1, 0, 0), "from": 0.2, "to": 2.0, "points": 4}])
    p = host.Params().set("scattering.average.orientation.type", "vectors")
    p.set("scattering.average.orientation.vectors.resolution", 20).set("scattering.average.orientation.vectors.seed", 5)
    p.create()
    recs, has, tm = host.run_scatter(p, xyz, qv, factors_fn=lambda ql: b * (1.0 + 0.1 * ql))
    assert has and len(recs) == 4 and tm["sd:compute"][1] == 1 and tm["sd:c:scan"][1] == 1  # one batched |q| scan
    for r, q in zip(recs, qv):
        ref = oracle.compute_all_vectors(xyz, b * (1.0 + 0.1 * np.linalg.norm(q)), p.init_subvectors(q), nthreads=4)
        assert rel_err(r["fqt"], ref[0]) < TOL and abs(r["fq"] - ref[1]) < TOL * abs(ref[0][0])
        assert r["fq0"] == r["fqt"][0]
    # cylinder orientation vectors
    p.set("scattering.average.orientation.vectors.type", "cylinder").set("scattering.average.orientation.axis.x", 1).create()
    recs, _, _ = host.run_scatter(p, xyz, [[0.3, 0.2, 0.5]], b=b)
    ref = oracle.compute_all_vectors(xyz, b, p.init_subvectors([0.3, 0.2, 0.5]), nthreads=4)
    assert rel_err(recs[0]["fqt"], ref[0]) < TOL
    # self
    ps = host.Params().set("scattering.type", "self").set("scattering.average.orientation.type", "vectors")
    ps.set("scattering.average.orientation.vectors.resolution", 6).create()
    recs, _, _ = host.run_scatter(ps, xyz[:, :40], qv[:2], b=b[:40])
    for r, q in zip(recs, qv[:2]):
        ref = oracle.compute_self_vectors(xyz[:, :40].transpose(1, 0, 2), b[:40], ps.init_subvectors(q), nthreads=4)
        assert rel_err(r["fqt"], ref[0]) < TOL
    # multipole cylinder (MPCylinderScatterDevice) around a tilted axis
    pc = host.Params().set("scattering.average.orientation.type", "multipole")
    pc.set("scattering.average.orientation.multipole.type", "cylinder").set("scattering.average.orientation.axis.y", 1)
    pc.set("scattering.average.orientation.multipole.moments.type", "resolution")
    pc.set("scattering.average.orientation.multipole.moments.resolution", 4).create()
    qc = np.array([[0.3, 0.2, 0.5], [-0.4, 0.0, 0.1]])
    recs, _, _ = host.run_scatter(pc, xyz, qc, b=b)
    for r, q in zip(recs, qc):
        ref = oracle.compute_mpcylinder(oracle.cart_to_cylindrical(xyz, (0, 1, 1)), b, q, (0, 1, 1), pc.moments, nthreads=4)
        assert rel_err(r["fqt"], ref[0]) < TOL and abs(r["fq"] - ref[1]) < TOL * abs(ref[0][0])
    # multipole sphere
    pm = host.Params().set("scattering.average.orientation.type", "multipole")
    pm.set("scattering.average.orientation.multipole.moments.type", "resolution")
    pm.set("scattering.average.orientation.multipole.moments.resolution", 5).create()
    recs, _, _ = host.run_scatter(pm, xyz, qv[1:3], b=b)
    for r, q in zip(recs, qv[1:3]):
        ref = oracle.compute_mpsphere(oracle.cart_to_spherical(xyz), b, np.linalg.norm(q), pm.moments, nthreads=4)
        assert rel_err(r["fqt"], ref[0]) < TOL


@pytest.mark.parametrize("dsp", ["autocorrelate", "square"])
def test_self_streamed_waves_match_resident(oracle, dsp):
    """BASELINE config 5 shape in small: the rank's atoms exceed limits.stage.memory.data, so the stager streams them
    through the GPU in waves (sgpu_stage_atoms_wave), every wave is evaluated for all |q| and the packed partials add up
    (sgpu_accumulate).  Same fqt / fq / fq2 as the resident run (1e-12) and as the oracle (1e-9)."""
    from sassena_b200 import host
    NA, NF = 61, 130
    xyz = synth.trajectory(NF, NA, 30.0, 0.2, 41, offset=-15.0)
    b = synth.factors(NA)
    qv = host.create_from_scans([{"base": (0, 1, 0), "from": 0.4, "to": 1.6, "points": 3}])
    ps = host.Params().set("scattering.type", "self").set("scattering.average.orientation.type", "vectors")
    ps.set("scattering.average.orientation.vectors.resolution", 7).set("scattering.dsp.type", dsp).create()
    resident, _, _ = host.run_scatter(ps, xyz, qv, factors_fn=lambda ql: b * (1.0 + 0.2 * ql))
    ps.set("limits.stage.memory.data", 32 * NF * 12)  # two wave buffers of 16 atoms -> 4 waves (16, 16, 16, 13)
    streamed, _, tm = host.run_scatter(ps, xyz, qv, factors_fn=lambda ql: b * (1.0 + 0.2 * ql))
    assert tm["sd:compute"][1] == 4 and len(streamed) == 3
    for r, s, q in zip(resident, streamed, qv):
        ref = oracle.compute_self_vectors(xyz.transpose(1, 0, 2), b * (1.0 + 0.2 * np.linalg.norm(q)), ps.init_subvectors(q),
                                          dsp=dsp, nthreads=4)
        assert rel_err(s["fqt"], r["fqt"]) < 1e-12 and abs(s["fq2"] - r["fq2"]) <= 1e-12 * abs(r["fq2"])
        assert rel_err(s["fqt"], ref[0]) < TOL and abs(s["fq"] - ref[1]) < TOL * abs(ref[0][0])
        assert abs(s["fq2"] - ref[2]) <= TOL * abs(ref[2])


def test_stage_atoms_wave_and_accumulate(gpu_ctx, oracle):
    """sgpu_stage_atoms_wave stages exactly atoms first + i*stride; sgpu_accumulate adds packed partials; bad ranges fail."""
    NA, NF, NM = 37, 40, 4
    frames = synth.trajectory(NF, NA, 30.0, 0.1, 19)
    b = synth.factors(NA)
    q = 0.9 * synth.unit_vectors(NM, 3)
    ids = 2 + 3 * np.arange(11)  # atoms 2, 5, ..., 32
    gpu_ctx.stage_atoms_wave(frames, 2, 3, 11)
    gpu_ctx.set_factors(b[ids])
    fqt, fq, fq2 = gpu_ctx.compute_self_vectors(q)
    rfqt, rfq, rfq2 = oracle.compute_self_vectors(np.ascontiguousarray(frames[:, ids].transpose(1, 0, 2)), b[ids], q)
    assert rel_err(fqt, rfqt) < TOL and abs(fq2 - rfq2) <= TOL * abs(rfq2)
    plen = gpu_ctx.partial_len("autocorrelate")
    d1, d2 = gpu_ctx.device_alloc(plen * 8), gpu_ctx.device_alloc(plen * 8)
    gpu_ctx.compute_self_vectors_partial(q, d1)
    gpu_ctx.compute_self_vectors_partial(q, d2)
    gpu_ctx.accumulate(d1, d2, plen)
    gpu_ctx.synchronize()
    a, c = np.empty(plen), np.empty(plen)
    gpu_ctx.memcpy_d2h(a, d1)
    gpu_ctx.memcpy_d2h(c, d2)
    assert np.array_equal(a, 2 * c)
    gpu_ctx.device_free(d1)
    gpu_ctx.device_free(d2)
    with pytest.raises(Exception, match="outside the trajectory"):
        gpu_ctx.stage_atoms_wave(frames, 2, 3, 13)  # 2 + 12*3 = 38 >= NA


@pytest.mark.parametrize("axis,q,L,NA", [((0, 0, 1), (0.3, -0.4, 0.5), 4, 300), ((1, 1, 0), (-0.7, 0.2, 0.1), 10, 517),
                                         ((0, 1, 0), (0.0, 0.0, 0.9), 2, 64), ((2, -1, 3), (1.5, 1.0, -2.0), 20, 1000),
                                         ((0, 0, 1), (-0.05, -0.05, 0.0), 6, 129), ((0, 0, 1), (0.0, 0.9, 0.2), 0, 31)])
def test_mpcylinder_matches_oracle(gpu_ctx, oracle, axis, q, L, NA):
    """K6 multipole cylinder: cartesian -> cylindrical conversion on the device (bit-exact with the oracle's float
    coordinates up to the last ulp of atan), amplitudes of all moments and the full fqt / fq / fq2 against the oracle;
    Bessel arguments from 0 to ~100 cover both the upward recurrence and the continued-fraction ratios."""
    NF = 37
    xyz = synth.trajectory(NF, NA, 60.0, 0.2, 23, offset=-30.0)
    xyz[0, 0] = 0.0  # an atom at the origin: r = 0, phi = 0
    xyz[1, 1, :2] = 0.0  # an atom on the z axis
    b = synth.factors(NA)
    mom = oracle.moments_cylinder(L)
    cyl = oracle.cart_to_cylindrical(xyz, axis)
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.frames_to_cylindrical(axis)
    gpu_ctx.set_factors(b)
    for dsp in ("autocorrelate", "square"):
        fqt, fq, fq2 = gpu_ctx.compute_mpcylinder(q, axis, mom, dsp=dsp)
        rfqt, rfq, rfq2 = oracle.compute_mpcylinder(cyl, b, q, axis, mom, dsp=dsp, nthreads=4)
        assert rel_err(fqt, rfqt) < TOL
        assert abs(fq - rfq) < TOL * abs(rfqt[0]) and abs(fq2 - rfq2) <= TOL * abs(rfq2)
    A = gpu_ctx.get_amplitudes(len(mom))
    *_, RA = oracle.compute_mpcylinder(cyl, b, q, axis, mom, dsp="plain", return_amplitudes=True, nthreads=4)
    assert np.max(np.abs(A - RA)) < TOL * np.max(np.abs(RA))


def test_mpcylinder_atom_sharding_and_errors(gpu_ctx, oracle):
    """atom ranges sum to the full amplitudes (multi-GPU decomposition of the device); the reference's moment checks"""
    NA, NF, axis, q = 333, 20, (1.0, 0.0, 1.0), (0.4, 0.3, -0.2)
    xyz = synth.trajectory(NF, NA, 40.0, 0.2, 29, offset=-20.0)
    b = synth.factors(NA)
    mom = oracle.moments_cylinder(5)
    gpu_ctx.stage_frames(xyz)
    with pytest.raises(Exception, match="cylindrical representation"):
        gpu_ctx.set_factors(b)
        gpu_ctx.compute_mpcylinder(q, axis, mom)
    gpu_ctx.frames_to_cylindrical(axis)
    gpu_ctx.set_factors(b)
    n = len(mom) * NF * 2
    d = [gpu_ctx.device_alloc(n * 8) for _ in range(3)]
    parts = []
    for k, (a0, na) in enumerate([(0, 100), (100, 133), (233, 100)]):
        gpu_ctx.mpcylinder_amplitudes(q, axis, mom, a0, na, d[k])
        gpu_ctx.synchronize()
        h = np.empty(n)
        gpu_ctx.memcpy_d2h(h, d[k])
        parts.append(h)
    total = (parts[0] + parts[1] + parts[2]).view(np.complex128).reshape(len(mom), NF)
    *_, RA = oracle.compute_mpcylinder(oracle.cart_to_cylindrical(xyz, axis), b, q, axis, mom, dsp="plain",
                                       return_amplitudes=True, nthreads=4)
    assert np.max(np.abs(total - RA)) < TOL * np.max(np.abs(RA))
    for p in d:
        gpu_ctx.device_free(p)
    with pytest.raises(Exception, match="between 0 and 3"):
        gpu_ctx.compute_mpcylinder(q, axis, [[1, 4]])
    with pytest.raises(Exception, match="must be 0 for Major 0"):
        gpu_ctx.compute_mpcylinder(q, axis, [[0, 2]])
    with pytest.raises(Exception, match="different axis"):
        gpu_ctx.compute_mpcylinder(q, (0, 0, 1), mom)


@pytest.mark.parametrize("L", [0, 1, 6, 20])
def test_mpsphere_batch_and_atom_sharding(gpu_ctx, oracle, L):
    """batched multipole sphere (several |q| per pass, per-|q| factors) and the atom-sharded multi-GPU protocol:
    amplitudes of atom slices summed == full amplitudes; everything vs the oracle"""
    NA, NF = 137, 9
    xyz = synth.trajectory(NF, NA, 60.0, 0.3, 23, offset=-30.0)
    sph = oracle.cart_to_spherical(xyz)
    b = synth.factors(NA)
    mom = oracle.moments_sphere(L)
    qls = np.array([0.01, 0.07, 0.3, 0.9, 1.7, 2.5, 0.5, 0.11, 1.23, 0.02, 3.1])  # 11 = 8 + 2 + 1
    bq = np.array([b * (1.0 + 0.05 * i) for i in range(len(qls))])
    from sassena_b200 import REPR_SPHERICAL
    gpu_ctx.stage_frames(sph, repr=REPR_SPHERICAL)
    gpu_ctx.set_factors(b)
    gpu_ctx.set_factors_batch(bq)
    res = gpu_ctx.compute_mpsphere_batch(qls, mom, dsp="autocorrelate")
    for i, ql in enumerate(qls):
        rfqt, rfq, rfq2 = oracle.compute_mpsphere(sph, bq[i], ql, mom, nthreads=4)
        assert rel_err(res[i][0], rfqt) < TOL
        assert abs(res[i][1] - rfq) < TOL * abs(rfqt[0])
        assert abs(res[i][2] - rfq2) <= TOL * abs(rfq2)
    # single-|q| entry point ignores the batch factors
    fqt, fq, fq2 = gpu_ctx.compute_mpsphere(qls[3], mom)
    rfqt, _, _ = oracle.compute_mpsphere(sph, b, qls[3], mom, nthreads=4)
    assert rel_err(fqt, rfqt) < TOL
    # atom sharding over 3 logical ranks (DivAssignment over atoms) + summed amplitudes
    NQ, NM = 3, len(mom)
    n = NQ * NM * NF * 2
    d_amp = gpu_ctx.device_alloc(n * 8)
    total = np.zeros(n)
    for r in range(3):
        off, size, _ = oracle.div_assignment(3, r, NA)
        gpu_ctx.mpsphere_amplitudes(qls[:NQ], mom, off, size, d_amp)
        gpu_ctx.synchronize()
        part = np.empty(n)
        gpu_ctx.memcpy_d2h(part, d_amp)
        total += part
    gpu_ctx.memcpy_h2d(d_amp, total)
    plen = gpu_ctx.partial_len("square")
    d_part = gpu_ctx.device_alloc(NQ * plen * 8)
    gpu_ctx.mpsphere_dsp_partial(d_amp, NQ, NM, d_part, dsp="square")
    for i in range(NQ):
        fqt, fq, fq2 = gpu_ctx.finalize(d_part + i * plen * 8, 1.0 / (4 * np.pi), dsp="square")
        rfqt, rfq, rfq2 = oracle.compute_mpsphere(sph, b, qls[i], mom, dsp="square", nthreads=4)
        assert rel_err(fqt, rfqt) < TOL and abs(fq - rfq) < TOL * abs(rfqt[0])
    gpu_ctx.device_free(d_amp)
    gpu_ctx.device_free(d_part)


def test_mpsphere_large_l_falls_back(gpu_ctx, oracle):
    """l > 21 has more (l,m) pairs than the batched kernel has threads: the per-|q| kernel takes over"""
    NA, NF = 40, 3
    xyz = synth.trajectory(NF, NA, 40.0, 0.3, 29, offset=-20.0)
    sph = oracle.cart_to_spherical(xyz)
    b = synth.factors(NA)
    mom = np.array([[24, -3], [24, 24], [30, 0], [2, 1]])
    from sassena_b200 import REPR_SPHERICAL
    gpu_ctx.stage_frames(sph, repr=REPR_SPHERICAL)
    gpu_ctx.set_factors(b)
    res = gpu_ctx.compute_mpsphere_batch([0.4, 1.1], mom, dsp="square")
    for i, ql in enumerate([0.4, 1.1]):
        rfqt, rfq, rfq2 = oracle.compute_mpsphere(sph, b, ql, mom, dsp="square")
        assert rel_err(res[i][0], rfqt) < TOL


def test_host_layer_mpsphere_batches_q(oracle):
    """MPSphereScatterDevice::runner batches |q| values (8 per pass) with per-|q| factors: 11 |q| = 8 + 3"""
    from sassena_b200 import host
    NA, NF = 90, 12
    xyz = synth.trajectory(NF, NA, 50.0, 0.2, 37, offset=-25.0)
    b = synth.factors(NA)
    qv = host.create_from_scans([{"base": (0, 1, 0), "from": 0.05, "to": 2.2, "points": 11}])
    pm = host.Params().set("scattering.average.orientation.type", "multipole")
    pm.set("scattering.average.orientation.multipole.moments.type", "resolution")
    pm.set("scattering.average.orientation.multipole.moments.resolution", 4).create()
    recs, has, tm = host.run_scatter(pm, xyz, qv, factors_fn=lambda ql: b * (1.0 + ql))
    assert len(recs) == 11 and tm["sd:compute"][1] == 2 and tm["sd:write"][1] == 11
    sph = oracle.cart_to_spherical(xyz)
    for r, q in zip(recs, qv):
        ql = np.linalg.norm(q)
        ref = oracle.compute_mpsphere(sph, b * (1.0 + ql), ql, pm.moments, nthreads=4)
        assert np.array_equal(r["q"], q)
        assert rel_err(r["fqt"], ref[0]) < TOL and abs(r["fq"] - ref[1]) < TOL * abs(ref[0][0])


@pytest.mark.parametrize("NM", [1, 5, 47, 48, 49, 62, 63, 96, 125, 143])
def test_all_vectors_exact_tail_launch(gpu_ctx, oracle, NM):
    """q-vector counts that are not multiples of the 48-vector CTA: the remainder goes through an exactly-sized tail
    launch (QPT 4..6, 1..8 warps); counts typical for 2/4/8-GPU sharding of 500 vectors (250, 125, 62/63) included"""
    NA, NF = 257, 6
    xyz = synth.trajectory(NF, NA, 25.0, 0.3, 41)
    b = synth.factors(NA)
    q = 1.9 * synth.unit_vectors(NM, 42)
    gpu_ctx.stage_frames(xyz)
    gpu_ctx.set_factors(b)
    gpu_ctx.compute_all_vectors(q, dsp="plain")
    A = gpu_ctx.get_amplitudes(NM)
    *_, Aref = oracle.compute_all_vectors(xyz, b, q, dsp="plain", return_amplitudes=True)
    assert rel_err(A, Aref) < 1e-12
