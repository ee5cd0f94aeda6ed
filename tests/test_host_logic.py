"""Host layer (C++ mirror of scatter_devices / stager / decomposition, csrc/host) against the oracle restatement and
the reference's documented behaviour.  CPU only: no compute calls into the CUDA library."""
import ctypes as C
import itertools
import os
import re

import numpy as np
import pytest

from sassena_b200 import _lib, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """the C-ABI library loads and exports every symbol include/*.h declares"""
    lib = _lib.load_library()
    for hdr in ("sassena_b200.h", "sassena_host.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        names = set(re.findall(r"\b(sgpu_[a-z0-9_]+|sass_[a-z0-9_]+)\s*\(", text))
        names -= {"sass_factors_fn", "sass_write_fn"}
        assert len(names) > 10
        for n in sorted(names):
            assert hasattr(lib, n), f"{n} declared in {hdr} but not exported"
    assert lib.sgpu_version().startswith(b"sassena_b200")


def test_no_cpu_fallback_without_gpu():
    """the product fails loudly when there is no CUDA device: no CPU path"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sassena_b200
    with pytest.raises(sassena_b200.SgpuError) as e:
        sassena_b200.ScatterContext(0)
    assert "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("NN,NAF", list(itertools.product([1, 2, 3, 4, 7, 8, 16], [1, 2, 5, 10, 100, 1001])))
def test_assignments_match_oracle_and_partition(oracle, NN, NAF):
    """DivAssignment / ModAssignment (assignment.cpp:27-132): same arithmetic as the oracle, and the ranks' index
    sets partition [0, NAF)"""
    seen_div, seen_mod = [], []
    for r in range(NN):
        assert host.div_assignment(NN, r, NAF) == oracle.div_assignment(NN, r, NAF)
        assert host.mod_assignment(NN, r, NAF) == oracle.mod_assignment(NN, r, NAF)
        off, size, mx = host.div_assignment(NN, r, NAF)
        seen_div += list(range(off, off + size))
        assert size <= mx == -(-NAF // NN)
        off, size, mx = host.mod_assignment(NN, r, NAF)
        seen_mod += [off + i * NN for i in range(size)]
    assert seen_div == list(range(NAF))
    assert sorted(seen_mod) == list(range(NAF))


def test_decomposition_plan_worked_example():
    """config 1 under 4 ranks (NQ=10, NAF=NF=100): ties go to the largest partition"""
    assert host.decomposition_plan(4, 10, 100, 1000 * 12, 500 << 20) == (1, 4, 0)
    # manual partition size 2 -> 2 partitions x 2
    assert host.decomposition_plan(4, 10, 100, 12000, 500 << 20, automatic=False, manual_size=2)[:2] == (2, 2)
    # memory limit forces more ranks per partition: NAFcycles*elbytes <= limit
    p, ps, _ = host.decomposition_plan(8, 50, 10000, 1200000, 1250 * 1200000)
    assert ps == 8 and p == 1
    with pytest.raises(host.HostError):
        host.decomposition_plan(8, 50, 10000, 1200000, 100 * 1200000)  # nothing fits
    with pytest.raises(host.HostError):
        host.decomposition_plan(7, 3, 5, 10, 1 << 30, utilization=0.99)  # utilisation too low


@pytest.mark.parametrize("nn,nq,naf", [(4, 10, 100), (8, 50, 10000), (8, 20, 30000), (3, 7, 11), (16, 5, 3), (6, 1, 1000),
                                       (5, 200, 1000)])
def test_decomposition_plan_matches_oracle(oracle, nn, nq, naf):
    rc, p, ps, pen = oracle.decomposition_plan(nn, nq, naf, 1200, 1 << 40, 0.0)
    assert rc == 0
    assert host.decomposition_plan(nn, nq, naf, 1200, 1 << 40, utilization=0.0) == (p, ps, pen)


def test_scans_use_float_rounded_fractions(oracle):
    """create_from_scans (parameters.cpp:1125-1189): powf() fractions, 1/2/3 nested scans"""
    s1 = [{"base": (1, 0, 0), "from": 0.2, "to": 2.0, "points": 10}]
    q = host.create_from_scans(s1)
    assert np.array_equal(q, oracle.qvectors_from_scans(s1))
    frac = np.float32(np.float32(3.0 / 9.0) ** np.float32(1.0))
    assert q[3, 0] == 0.2 + float(frac) * (2.0 - 0.2)
    assert q[3, 0] != 0.2 + (3.0 / 9.0) * 1.8  # not the double-precision fraction
    s3 = [{"base": (1, 0, 0), "from": 0, "to": 1, "points": 3}, {"base": (0, 1, 0), "from": 0, "to": 2, "points": 2},
          {"base": (0, 0, 1), "from": 1, "to": 3, "points": 1, "exponent": 2.0}]
    q3 = host.create_from_scans(s3)
    assert q3.shape == (6, 3)
    assert np.array_equal(q3, oracle.qvectors_from_scans(s3))
    with pytest.raises(host.HostError):
        host.create_from_scans(s3 + s1)


def test_orientation_generators(oracle):
    """parameters.cpp:930-1034: sphere / cylinder (boost_uniform_on_sphere, raster_linear) / file"""
    p = host.Params().set("scattering.average.orientation.type", "vectors")
    p.set("scattering.average.orientation.vectors.resolution", 50).set("scattering.average.orientation.vectors.seed", 7)
    p.create()
    v = p.vectors
    assert v.shape == (50, 3)
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-15)
    assert np.array_equal(v, oracle.uniform_on_sphere(7, 3, 50))
    p.set("scattering.average.orientation.vectors.type", "cylinder").create()
    c = p.vectors
    assert np.array_equal(c, oracle.uniform_on_sphere(7, 2, 50)) and np.all(c[:, 2] == 0)
    p.set("scattering.average.orientation.vectors.algorithm", "raster_linear")
    p.set("scattering.average.orientation.vectors.resolution", 1).create()
    r = p.vectors
    assert np.array_equal(r, oracle.cylinder_raster_linear(1)) and len(r) in (360, 361)
    p.set("scattering.average.orientation.vectors.type", "file").set_vectors([[2, 0, 0], [0, 0, 0], [1, 1, 1]]).create()
    f = p.vectors
    assert np.allclose(f, [[1, 0, 0], [0, 0, 0], np.ones(3) / np.sqrt(3)])
    p.set("scattering.average.orientation.vectors.type", "sphere")
    p.set("scattering.average.orientation.vectors.algorithm", "quad")
    with pytest.raises(host.HostError):
        p.create()


def test_mt19937_reference_stream(oracle):
    """boost::mt19937 default-seeded 10000th output is 4123659995 (the generator's published check value)"""
    s = oracle.mt19937_stream(5489, 10000)
    assert int(s[-1]) == 4123659995


def test_moments_generator(oracle):
    p = host.Params().set("scattering.average.orientation.type", "multipole")
    with pytest.raises(host.HostError):
        p.create()  # moments.type has no default in the reference (parameters.cpp:1037-1079)
    p.set("scattering.average.orientation.multipole.moments.type", "resolution")
    p.set("scattering.average.orientation.multipole.moments.resolution", 20).create()
    m = p.moments
    assert len(m) == 441 and np.array_equal(m, oracle.moments_sphere(20))
    assert tuple(m[0]) == (0, 0) and tuple(m[1]) == (1, -1) and tuple(m[-1]) == (20, 20)
    p.set("scattering.average.orientation.multipole.moments.type", "file").set_moments([[1, 2]])
    with pytest.raises(host.HostError):
        p.create()  # |m| > l


@pytest.mark.parametrize("axis,q", [((0, 0, 1), (0.3, 0.0, 0.4)), ((1, 1, 0), (0.1, 0.2, 0.3)), ((0, 0, 1), (0, 0, 0.7)),
                                    ((0, 1, 0), (1.0, 0, 0))])
def test_init_subvectors(oracle, axis, q):
    """abstract_vectors_scatter_device.cpp:112-175 for none / sphere / cylinder (incl. the q_r == 0 quirk)"""
    p = host.Params()
    assert np.array_equal(p.init_subvectors(q), [q])
    p.set("scattering.average.orientation.type", "vectors").set("scattering.average.orientation.vectors.resolution", 9)
    p.create()
    u = p.vectors
    assert np.array_equal(p.init_subvectors(q), oracle.init_subvectors("sphere", q, u))
    for k, a in zip("xyz", axis):
        p.set(f"scattering.average.orientation.axis.{k}", a)
    p.set("scattering.average.orientation.vectors.type", "cylinder").create()
    c = p.vectors
    got = p.init_subvectors(q)
    assert np.array_equal(got, oracle.init_subvectors("cylinder", q, c, axis))
    base = oracle.vector_base(axis)
    if np.hypot(*(base[:2] @ np.array(q))) == 0:
        assert len(got) == 1
    else:
        # every cylinder subvector keeps the component along the axis and the radial length
        ez = base[2]
        assert np.allclose(got @ ez, np.dot(q, ez))
        assert np.allclose(np.linalg.norm(got - np.outer(got @ ez, ez), axis=1), np.linalg.norm(q - np.dot(q, ez) * ez))


def test_params_errors():
    p = host.Params()
    with pytest.raises(host.HostError) as e:
        p.set("scattering.target", "system")
    assert "obsolete" in str(e.value)
    with pytest.raises(host.HostError):
        p.set("no.such.key", 1)
    p.set("limits.decomposition.partitions.automatic", "TRUE").set("limits.decomposition.partitions.automatic", "0")
    with pytest.raises(host.HostError):
        p.set("limits.decomposition.partitions.automatic", "maybe")


def test_run_scatter_single_rank_oracle_backend(oracle):
    """factory + AllVectors/SelfVectors/MPSphere device.run() over the oracle-bound backend table: the host control
    flow (init_subvectors, factors per |q|, write per |q|, 1/NM and 1/4pi scaling) reproduces the oracle."""
    from oracle_backend import OracleBackend
    from sassena_b200 import synth
    be = OracleBackend()
    NA, NF = 30, 16
    xyz = synth.trajectory(NF, NA, 20.0, 0.2, 3, offset=-10.0)
    b = synth.factors(NA)
    qv = host.create_from_scans([{"base": (1, 0, 0), "from": 0.3, "to": 1.5, "points": 3}])
    # coherent, sphere vectors given as file rows; |q|-dependent factors through the callback
    p = host.Params().set("scattering.average.orientation.type", "vectors")
    p.set("scattering.average.orientation.vectors.type", "file").set_vectors(synth.unit_vectors(6, 1)).create()
    recs, has, tm = host.run_scatter(p, xyz, qv, factors_fn=lambda ql: b * (1 + ql), backend=be.vtbl)
    assert has and len(recs) == 3 and tm["sd:compute"][1] == 3 and "sd:stage" in tm
    for r, q in zip(recs, qv):
        ref = oracle.compute_all_vectors(xyz, b * (1 + np.linalg.norm(q)), p.init_subvectors(q))
        assert np.allclose(r["fqt"], ref[0], rtol=1e-12, atol=1e-12 * abs(ref[0][0]))
        assert np.isclose(r["fq"], ref[1]) and np.isclose(r["fq2"], ref[2]) and r["fq0"] == r["fqt"][0]
        assert np.array_equal(r["q"], q)
    # self, no orientational average, direct method
    p2 = host.Params().set("scattering.type", "self").set("scattering.dsp.method", "direct")
    recs, _, _ = host.run_scatter(p2, xyz, qv[:1], b=b, backend=be.vtbl)
    ref = oracle.compute_self_vectors(xyz.transpose(1, 0, 2), b, qv[:1], method="direct")
    assert np.allclose(recs[0]["fqt"], ref[0], rtol=1e-11, atol=1e-11 * abs(ref[0][0]))
    # self, streamed: the atoms do not fit limits.stage.memory.data and pass through the device in waves (BASELINE
    # config 5); same result as the resident run, and the reference's error when streaming is switched off
    p2s = host.Params().set("scattering.type", "self").set("scattering.average.orientation.type", "vectors")
    p2s.set("scattering.average.orientation.vectors.type", "file").set_vectors(synth.unit_vectors(4, 5)).create()
    resident, _, _ = host.run_scatter(p2s, xyz, qv, factors_fn=lambda ql: b * (1 + ql), backend=be.vtbl)
    be.waves_staged = 0
    p2s.set("limits.stage.memory.data", 14 * NF * 12)  # two wave buffers of 7 of the 30 atoms -> 5 waves
    streamed, _, tms = host.run_scatter(p2s, xyz, qv, factors_fn=lambda ql: b * (1 + ql), backend=be.vtbl)
    assert be.waves_staged == 5 and tms["sd:compute"][1] == 5 and len(streamed) == len(qv)
    for r, s, q in zip(resident, streamed, qv):
        ref = oracle.compute_self_vectors(xyz.transpose(1, 0, 2), b * (1 + np.linalg.norm(q)), p2s.init_subvectors(q))
        assert np.array_equal(s["q"], q)
        assert np.allclose(s["fqt"], r["fqt"], rtol=1e-12, atol=1e-12 * abs(r["fqt"][0]))
        assert np.allclose(s["fqt"], ref[0], rtol=1e-11, atol=1e-11 * abs(ref[0][0]))
        assert np.isclose(s["fq"], ref[1]) and np.isclose(s["fq2"], ref[2])
    with pytest.raises(host.HostError) as e:
        host.run_scatter(p2s.set("limits.stage.stream", False), xyz, qv, b=b, backend=be.vtbl)
    assert "decomposition failed" in str(e.value) or "Insufficient Buffer" in str(e.value)
    # multipole sphere
    p3 = host.Params().set("scattering.average.orientation.type", "multipole")
    p3.set("scattering.average.orientation.multipole.moments.type", "resolution")
    p3.set("scattering.average.orientation.multipole.moments.resolution", 3).set("scattering.dsp.type", "square").create()
    recs, _, _ = host.run_scatter(p3, xyz, qv[1:2], b=b, backend=be.vtbl)
    ref = oracle.compute_mpsphere(oracle.cart_to_spherical(xyz), b, np.linalg.norm(qv[1]), p3.moments, dsp="square")
    assert np.allclose(recs[0]["fqt"], ref[0], rtol=1e-12)
    # error sites of the reference
    with pytest.raises(host.HostError) as e:
        host.run_scatter(host.Params().set("scattering.dsp.type", "cube"), xyz, qv, b=b, backend=be.vtbl)
    assert "DSP type not understood" in str(e.value)
    with pytest.raises(host.HostError) as e:
        host.run_scatter(host.Params().set("scattering.type", "both"), xyz, qv, b=b, backend=be.vtbl)
    assert "Must be 'self' or 'all'" in str(e.value)
    with pytest.raises(host.HostError) as e:
        host.run_scatter(host.Params(), xyz, np.zeros((0, 3)), b=b, backend=be.vtbl)
    assert "No qvectors left to compute" in str(e.value)
    # multipole cylinder around a tilted axis (MPCylinderScatterDevice): 1 + 4*3 moments, result scaled by 1/(2 pi)
    pc = host.Params().set("scattering.average.orientation.type", "multipole")
    pc.set("scattering.average.orientation.multipole.type", "cylinder")
    pc.set("scattering.average.orientation.axis.x", 1).set("scattering.average.orientation.axis.y", 1)
    pc.set("scattering.average.orientation.axis.z", 0.5)
    pc.set("scattering.average.orientation.multipole.moments.type", "resolution")
    pc.set("scattering.average.orientation.multipole.moments.resolution", 3).create()
    assert len(pc.moments) == 13
    qc = np.array([[0.3, -0.2, 0.6], [-0.5, 0.1, 0.0]])
    recs, _, _ = host.run_scatter(pc, xyz, qc, b=b, backend=be.vtbl)
    for r, q in zip(recs, qc):
        ref = oracle.compute_mpcylinder(oracle.cart_to_cylindrical(xyz, (1, 1, 0.5)), b, q, (1, 1, 0.5), pc.moments)
        assert np.allclose(r["fqt"], ref[0], rtol=1e-12, atol=1e-12 * abs(ref[0][0]))
        assert np.isclose(r["fq"], ref[1]) and np.isclose(r["fq2"], ref[2])
    with pytest.raises(host.HostError) as e:
        host.run_scatter(host.Params().set("limits.stage.memory.data", 100), xyz, qv, b=b, backend=be.vtbl)
    assert "decomposition failed" in str(e.value) or "Insufficient Buffer" in str(e.value)


def test_run_scatter_scan_batching(oracle):
    """AllVectors runner: consecutive q-vectors that share their directions (>= 4 of them, >= 8 subvectors) go to the
    backend in one call, whatever their spacing; a different direction or |q| = 0 ends the batch; the records still
    arrive one per q-vector, in order, and equal the oracle.  limits.computation.scan=1 disables it."""
    from oracle_backend import OracleBackend
    from sassena_b200 import synth
    be = OracleBackend()
    NA, NF = 20, 8
    xyz = synth.trajectory(NF, NA, 20.0, 0.2, 3, offset=-10.0)
    b = synth.factors(NA)
    scan = host.create_from_scans([{"base": (1, 2, 0), "from": 0.3, "to": 1.5, "points": 5}])
    qv = np.concatenate([scan, [[0.0, 0.0, 2.5]], scan[::-1][:2] * 1.7, [[0.0, 0.0, 0.0]]])  # 5 + 1 + 2 + 1 vectors

    def params(**kv):
        p = host.Params().set("scattering.average.orientation.type", "vectors")
        p.set("scattering.average.orientation.vectors.type", "file").set_vectors(synth.unit_vectors(9, 1))
        for k, v in kv.items():
            p.set(k.replace("__", "."), v)
        return p.create()

    p = params()
    recs, has, tm = host.run_scatter(p, xyz, qv, factors_fn=lambda ql: b * (1 + ql), backend=be.vtbl)
    assert len(recs) == len(qv)
    # sphere/file averaging only depends on |q|: all 8 non-zero vectors share their subvector directions and form one
    # batch; |q| = 0 never qualifies
    assert tm["sd:c:scan"][1] == 1 and tm["sd:compute"][1] == 1 + 1
    for r, q in zip(recs, qv):
        assert np.array_equal(r["q"], q)
        ref = oracle.compute_all_vectors(xyz, b * (1 + np.linalg.norm(q)), p.init_subvectors(q))
        assert np.allclose(r["fqt"], ref[0], rtol=1e-11, atol=1e-11 * abs(ref[0][0]))
        assert np.isclose(r["fq"], ref[1], rtol=1e-11, atol=1e-11 * abs(ref[0][0]))
    # cylinder averaging is linear in |q| for a fixed direction: batched too
    pc = params(scattering__average__orientation__vectors__type="cylinder", scattering__average__orientation__vectors__resolution=11)
    recs, _, tm = host.run_scatter(pc, xyz, scan, b=b, backend=be.vtbl)
    assert tm["sd:c:scan"][1] == 1
    # ... but not across directions: the cylinder construction depends on the direction of q
    _, _, tm2 = host.run_scatter(pc, xyz, qv[:8], b=b, backend=be.vtbl)
    assert tm2["sd:c:scan"][1] == 1 and tm2["sd:compute"][1] == 1 + 3
    for r, q in zip(recs, scan):
        ref = oracle.compute_all_vectors(xyz, b, pc.init_subvectors(q))
        assert np.allclose(r["fqt"], ref[0], rtol=1e-11, atol=1e-11 * abs(ref[0][0]))
    # scan_snap: a float-rounded 7-point scan is moved onto the exact progression and the snapped q-vectors are written
    rounded = host.create_from_scans([{"base": (0, 1, 0), "from": 0.1, "to": 1.3, "points": 7}])
    sr = np.linalg.norm(rounded, axis=1)
    assert np.max(np.abs(np.diff(sr, 2))) > 1e-10
    recs, _, _ = host.run_scatter(params(limits__computation__scan_snap=True), xyz, rounded, b=b, backend=be.vtbl)
    snapped = np.array([r["q"] for r in recs])
    assert np.max(np.abs(np.diff(np.linalg.norm(snapped, axis=1), 2))) < 1e-15
    assert np.allclose(snapped, rounded, rtol=1e-6) and not np.array_equal(snapped, rounded)
    ref = oracle.compute_all_vectors(xyz, b, p.init_subvectors(snapped[3]))
    assert np.allclose(recs[3]["fqt"], ref[0], rtol=1e-11, atol=1e-11 * abs(ref[0][0]))
    recs, _, _ = host.run_scatter(params(), xyz, rounded, b=b, backend=be.vtbl)
    assert np.array_equal(np.array([r["q"] for r in recs]), rounded)  # default: the reference's q-vectors, untouched
    # switched off
    recs, _, tm = host.run_scatter(params(limits__computation__scan=1), xyz, scan, b=b, backend=be.vtbl)
    assert "sd:c:scan" not in tm and tm["sd:compute"][1] == 5 and len(recs) == 5


def test_dcd_roundtrip_and_trimming(tmp_path):
    """DCD writer (coordinate_writer.cpp layout) -> reader (frames.cpp:272-436): exact round trip, header fields,
    first/last/stride trimming with the reference's absolute-index stride rule (frames.cpp:224-245)"""
    import struct
    from sassena_b200 import synth
    xyz = synth.trajectory(11, 7, 20.0, 0.3, 5)
    path = tmp_path / "traj.dcd"
    host.write_dcd(path, xyz)
    raw = open(path, "rb").read()
    assert struct.unpack("<i", raw[:4])[0] == 84 and raw[4:8] == b"CORD" and struct.unpack("<i", raw[88:92])[0] == 84
    assert struct.unpack("<i", raw[8:12])[0] == 11
    # 92 header + title (4+4+4) + natoms (4+4+4) ; per frame: cell (4+48+4) + 3*(4+4N+4)
    assert len(raw) == 92 + 12 + 12 + 11 * (56 + 3 * (8 + 4 * 7))
    d = host.DCDFile(path)
    assert (d.number_of_frames, d.number_of_atoms, d.has_unitcell) == (11, 7, True)
    assert np.array_equal(d.read(), xyz)
    assert np.array_equal(d.read(3, 2), xyz[3:5])
    with pytest.raises(host.HostError):
        d.read(10, 5)
    # first=3, last=9, stride=2 keeps absolute indices 4, 6, 8 (i % stride == 0, not (i-first) % stride)
    t = host.DCDFile(path, first=3, last=9, stride=2)
    assert t.number_of_frames == 3 and np.array_equal(t.read(), xyz[[4, 6, 8]])
    # a file that is not a DCD
    bad = tmp_path / "bad.dcd"
    bad.write_bytes(b"\0" * 200)
    with pytest.raises(host.HostError) as e:
        host.DCDFile(bad)
    assert "appears to not be a DCD file" in str(e.value)


def test_dcd_reader_handles_title_and_optional_blocks(tmp_path):
    """hand-built DCD without the unit-cell block, with an 80-byte title line and with the 4th (ext block 2) record"""
    import struct
    NF, NA = 3, 5
    rng = np.random.default_rng(0)
    xyz = rng.normal(size=(NF, NA, 3)).astype(np.float32)
    hdr = struct.pack("<i4siii24sfii", 84, b"CORD", NF, 0, 1, b"\0" * 24, 1.0, 0, 1) + b"\0" * 28 + struct.pack("<ii", 24, 84)
    assert len(hdr) == 92
    title = struct.pack("<ii", 84, 1) + b"T" * 80 + struct.pack("<i", 84)
    natoms = struct.pack("<iii", 4, NA, 4)
    body = b""
    for f in range(NF):
        for c in range(3):
            body += struct.pack("<i", 4 * NA) + xyz[f, :, c].tobytes() + struct.pack("<i", 4 * NA)
        body += struct.pack("<i", 4 * NA) + np.zeros(NA, np.float32).tobytes() + struct.pack("<i", 4 * NA)
    p = tmp_path / "ext2.dcd"
    p.write_bytes(hdr + title + natoms + body)
    d = host.DCDFile(p)
    assert (d.number_of_frames, d.number_of_atoms, d.has_unitcell) == (NF, NA, False)
    assert np.array_equal(d.read(), xyz)


def test_dcd_header_frame_count_is_checked_against_the_file(tmp_path):
    """The reference sizes its frame index by the header's count alone (frames.cpp:261-268); a count that is negative or
    larger than the file can hold (a truncated or corrupt trajectory) is refused when the file is opened."""
    import struct
    from sassena_b200 import synth
    xyz = synth.trajectory(5, 4, 20.0, 0.3, 6)
    path = tmp_path / "traj.dcd"
    host.write_dcd(path, xyz)
    raw = bytearray(open(path, "rb").read())
    for count in (6, 2**31 - 1, -1, -2**31):
        bad = tmp_path / "count.dcd"
        bad.write_bytes(bytes(raw[:8]) + struct.pack("<i", count) + bytes(raw[12:]))
        with pytest.raises(host.HostError, match="announces"):
            host.DCDFile(bad)
    cut = tmp_path / "cut.dcd"  # the last frame is incomplete
    cut.write_bytes(bytes(raw[:-10]))
    with pytest.raises(host.HostError, match="announces 5 frames, the file holds 4"):
        host.DCDFile(cut)
    fewer = tmp_path / "fewer.dcd"  # a header that announces fewer frames than the file holds is taken at its word
    fewer.write_bytes(bytes(raw[:8]) + struct.pack("<i", 3) + bytes(raw[12:]))
    d = host.DCDFile(fewer)
    assert d.number_of_frames == 3 and np.array_equal(d.read(), xyz[:3])
