"""sample.motions / sample.alignments ("next" row): csrc/host/coordinate_sets.cpp against restatements of
the reference written here from src/sample/coordinate_sets.cpp:245-353 (per-frame order: pre-alignments, motions,
post-alignments), motion_walker.cpp (walkers), center_of_mass.cpp (mass-weighted centres, Kabsch fit), with the
narrowing to float at the end (data_stager.cpp:111-113).  Deterministic walkers and translations must agree BIT FOR BIT
(same double operations in the same order); the rotational fit is compared with numpy's SVD Kabsch rotation."""
import math

import numpy as np
import pytest

from sassena_b200 import host
from test_control_plane import NAME2EL, make_case

MASS = {"hydrogen": 1.008, "carbon": 12.011, "oxygen": 15.999, "nitrogen": 14.007}
SEL = """<selections>
  <selection><type>range</type><name>front</name><from>0</from><to>9</to></selection>
  <selection><type>index</type><name>odd</name><index>1</index><index>3</index><index>5</index><index>7</index><index>11</index></selection>
</selections>"""


def masses(names):
    return np.array([MASS[NAME2EL[n]] for n in names])


def com(x, m, idx):
    """CenterOfMass: sequential mass-weighted sums, then the division (center_of_mass.cpp:186-211)"""
    mt = xt = yt = zt = 0.0
    for i in idx:
        mi = float(m[i])
        mt += mi
        xt += float(x[i, 0]) * mi
        yt += float(x[i, 1]) * mi
        zt += float(x[i, 2]) * mi
    return np.array([xt / mt, yt / mt, zt / mt])


def job_frames(tmp_path, sample_extra, NA=24, NF=12, stager=""):
    cfg, xyz, names = make_case(tmp_path, NA=NA, NF=NF, sample_extra=SEL + sample_extra, stager=stager,
                                scattering="<vectors><type>single</type><single><x>1</x><y>0</y><z>0</z></single></vectors>")
    job = host.Job(cfg)
    return job, job.frames(), xyz.astype(np.float64), names


def test_no_motion_is_identity(tmp_path):
    job, fr, xyz, _ = job_frames(tmp_path, "")
    assert np.array_equal(fr, xyz.astype(np.float32))


@pytest.mark.parametrize("kind", ["linear", "fixed", "oscillation"])
def test_deterministic_walkers_bit_exact(tmp_path, kind):
    d = (1.0, 2.0, -2.0)
    extra = f"""<motions><motion><type>{kind}</type><displace>0.37</displace><sampling>3</sampling><frequency>0.013</frequency>
      <selection>front</selection><direction><x>{d[0]}</x><y>{d[1]}</y><z>{d[2]}</z></direction></motion></motions>"""
    job, fr, xyz, names = job_frames(tmp_path, extra)
    m = masses(names)
    NF, NA = xyz.shape[:2]
    dl = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
    exp = xyz.copy()
    for f in range(NF):
        if kind == "linear":
            base = [(0.37 * 3) * c / dl for c in d]
            t = [float(f) * c for c in base]
        elif kind == "fixed":
            t = [0.37 * c / dl for c in d]
        else:
            base = [0.37 * c / dl for c in d]
            sn = math.sin(2 * math.pi * f * 0.013 * 3)
            t = [sn * c for c in base]
        ref = com(exp[f], m, range(NA))  # default reference: instant, selection "system" (parameters.cpp:249-254)
        for i in range(10):
            for c in range(3):
                exp[f, i, c] = ((exp[f, i, c] + -1.0 * ref[c]) + t[c]) + ref[c]
    assert np.array_equal(fr, exp.astype(np.float32))
    # the other atoms are untouched
    assert np.array_equal(fr[:, 10:], xyz[:, 10:].astype(np.float32))


def test_center_alignment_pre_and_post_bit_exact(tmp_path):
    extra = """<motions><motion><type>fixed</type><displace>2.5</displace><selection>odd</selection></motion></motions>
    <alignments>
      <alignment><type>center</type><selection>system</selection><order>pre</order>
         <reference><selection>front</selection></reference></alignment>
      <alignment><type>center</type><selection>front</selection><order>post</order></alignment>
    </alignments>"""
    job, fr, xyz, names = job_frames(tmp_path, extra)
    m = masses(names)
    NF, NA = xyz.shape[:2]
    odd = [1, 3, 5, 7, 11]
    exp = xyz.copy()
    for f in range(NF):
        o = com(exp[f], m, range(10))               # pre: centre of `front`, whole system moved
        exp[f] += -1.0 * o
        ref = com(exp[f], m, range(NA))             # motion about the instant system centre
        for i in odd:
            for c in range(3):
                exp[f, i, c] = ((exp[f, i, c] + -1.0 * ref[c]) + (2.5 if c == 0 else 0.0)) + ref[c]
        o = com(exp[f], m, range(10))               # post: reference selection defaults to the alignment's selection
        exp[f, :10] += -1.0 * o
    assert np.array_equal(fr, exp.astype(np.float32))


def test_fittrans_against_reference_frame(tmp_path):
    extra = """<alignments><alignment><type>fittrans</type><selection>front</selection>
         <reference><type>frame</type><frame>3</frame><selection>front</selection></reference></alignment></alignments>"""
    job, fr, xyz, names = job_frames(tmp_path, extra)
    m = masses(names)
    exp = xyz.copy()
    ref = com(xyz[3], m, range(10))
    for f in range(xyz.shape[0]):
        pos = com(exp[f], m, range(10))
        exp[f, :10] += ref - pos
    assert np.array_equal(fr, exp.astype(np.float32))
    # every frame's `front` centre now coincides with frame 3's
    for f in range(xyz.shape[0]):
        assert np.allclose(com(fr[f].astype(np.float64), m, range(10)), ref, atol=1e-5)


def kabsch(cur, ref, w):
    """optimal proper rotation R (R @ v ~ u) about the weighted centres"""
    pc = (cur * w[:, None]).sum(0) / w.sum()
    rc = (ref * w[:, None]).sum(0) / w.sum()
    K = np.einsum("i,ia,ib->ab", w, ref - rc, cur - pc)
    U, S, Vt = np.linalg.svd(K)
    P = np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))])
    return U @ P @ Vt, pc, rc


@pytest.mark.parametrize("kind", ["fitrottrans", "fitrot"])
def test_rotational_fit_matches_kabsch(tmp_path, kind):
    extra = f"""<alignments><alignment><type>{kind}</type><selection>system</selection><order>post</order>
         <reference><type>frame</type><frame>0</frame><selection>system</selection></reference></alignment></alignments>"""
    job, fr, xyz, names = job_frames(tmp_path, extra, NA=16, NF=8)
    m = masses(names)
    ref = xyz[0]
    for f in range(xyz.shape[0]):
        R, pc, rc = kabsch(xyz[f], ref, m)
        exp = (xyz[f] - pc) @ R.T + rc
        if kind == "fitrot":
            exp = exp + pc  # the reference adds the old centre back after the fit (coordinate_sets.cpp:282-286)
        assert np.allclose(fr[f], exp, atol=2e-5), f
    # frame 0 fitted onto itself is unchanged (fitrottrans)
    if kind == "fitrottrans":
        assert np.allclose(fr[0], xyz[0], atol=1e-5)


def test_rotational_fit_removes_a_rigid_rotation(tmp_path):
    """frames = one structure rigidly rotated and shifted: fitrottrans onto frame 0 recovers frame 0 (incl. a mirror-
    free proper rotation), the check Kabsch's determinant correction exists for"""
    import os
    NA, NF = 16, 6
    cfg, xyz, names = make_case(tmp_path, NA=NA, NF=NF, sample_extra=SEL + """<alignments><alignment><type>fitrottrans</type>
        <reference><type>frame</type><frame>0</frame></reference></alignment></alignments>""",
                                scattering="<vectors><type>single</type><single><x>1</x><y>0</y><z>0</z></single></vectors>")
    rng = np.random.default_rng(5)
    base = xyz[0].astype(np.float64)
    frames = [base]
    for f in range(1, NF):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        frames.append(base @ q.T + rng.normal(size=3) * 5)
    host.write_dcd(os.path.join(str(tmp_path), "traj.dcd"), np.array(frames, dtype=np.float32))
    fr = host.Job(cfg).frames()
    for f in range(NF):
        assert np.allclose(fr[f], base, atol=2e-4), f


def boost_normal_stream(oracle, seed):
    """boost::normal_distribution over mt19937, Boost 1.4x Box-Muller with one cached value (oracle's uniform_on_sphere uses it)"""
    raw = oracle.mt19937_stream(seed, 4096)
    pos = 0
    cached = None
    while True:
        if cached is None:
            r1 = raw[pos] / 4294967296.0
            r2 = raw[pos + 1] / 4294967296.0
            pos += 2
            rho = math.sqrt(-2.0 * math.log(1.0 - r2))
            cached = rho * math.sin(2 * math.pi * r1)
            yield rho * math.cos(2 * math.pi * r1)
        else:
            v, cached = cached, None
            yield v


def sphere_stream(normal):
    while True:
        v = [next(normal) for _ in range(3)]
        n = math.sqrt(sum(c * c for c in v))
        inv = 1.0 / n
        yield [c * inv for c in v]


@pytest.mark.parametrize("kind", ["randomwalk", "brownian", "localbrownian"])
def test_random_walkers_follow_the_boost_streams(tmp_path, oracle, kind):
    extra = f"""<motions><motion><type>{kind}</type><displace>0.8</displace><seed>7</seed><sampling>2</sampling>
       <radius>1.5</radius><selection>front</selection>
       <reference><type>frame</type><frame>0</frame><selection>front</selection></reference></motion></motions>"""
    job, fr, xyz, names = job_frames(tmp_path, extra)
    m = masses(names)
    NF = xyz.shape[0]
    if kind == "randomwalk":
        sph = sphere_stream(boost_normal_stream(oracle, 7))
    else:
        nrm = boost_normal_stream(oracle, 7)
        sph = sphere_stream(boost_normal_stream(oracle, 8))
    t = np.zeros(3)
    trans = []
    for f in range(NF):
        while True:
            if kind == "randomwalk":
                v = next(sph)
                next(sph)  # sampling - 1 discarded draws
                nt = t + 0.8 * np.array(v)
            else:
                nr = next(nrm)
                v = next(sph)
                next(nrm)
                next(sph)
                nt = t + (0.8 * nr) * np.array(v)
            if kind == "localbrownian" and math.sqrt(float(nt @ nt)) > 1.5:
                continue
            break
        t = nt
        trans.append(t)
    ref = com(xyz[0], m, range(10))  # stored reference: frame 0, selection front
    exp = xyz.copy()
    for f in range(NF):
        for c in range(3):
            exp[f, :10, c] = ((exp[f, :10, c] + -1.0 * ref[c]) + trans[f][c]) + ref[c]
    assert np.allclose(fr, exp, atol=1e-5)
    if kind == "localbrownian":
        assert all(np.linalg.norm(x) <= 1.5 for x in trans)


def test_rotational_brownian_is_a_rigid_rotation_about_the_centre(tmp_path):
    extra = """<motions><motion><type>rotationalbrownian</type><displace>5</displace><seed>3</seed></motion></motions>"""
    job, fr, xyz, names = job_frames(tmp_path, extra)
    m = masses(names)
    for f in range(xyz.shape[0]):
        c0 = com(xyz[f], m, range(xyz.shape[1]))
        c1 = com(fr[f].astype(np.float64), m, range(xyz.shape[1]))
        assert np.allclose(c0, c1, atol=1e-5)  # centre of mass stays
        d0 = np.linalg.norm(xyz[f] - c0, axis=1)
        d1 = np.linalg.norm(fr[f] - c1, axis=1)
        assert np.allclose(d0, d1, atol=1e-4)  # distances to it too
    assert not np.allclose(fr[5], xyz[5], atol=1e-3)  # but the frame did rotate
    # cumulative: the rotation of frame f is that of frame f-1 times a small one
    def rot(f):
        R, *_ = kabsch(xyz[f], fr[f].astype(np.float64), m)
        return R
    ang = [math.degrees(math.acos(max(-1, min(1, (np.trace(rot(f).T @ rot(f + 1)) - 1) / 2)))) for f in range(6)]
    assert all(0 < a < 40 for a in ang)


def test_reference_file_and_errors(tmp_path):
    import os
    extra = """<alignments><alignment><type>fittrans</type><selection>system</selection>
       <reference><type>file</type><file>ref.pdb</file><format>pdb</format><selection>system</selection></reference>
       </alignment></alignments>"""
    cfg, xyz, names = make_case(tmp_path, sample_extra=SEL + extra,
                                scattering="<vectors><type>single</type><single><x>1</x><y>0</y><z>0</z></single></vectors>")
    NA = xyz.shape[1]
    refc = np.arange(NA * 3, dtype=np.float64).reshape(NA, 3) * 0.25
    with open(os.path.join(str(tmp_path), "ref.pdb"), "w") as f:
        for i in range(NA):
            f.write("ATOM  %5d  CA  ALA A   1    %8.3f%8.3f%8.3f  1.00  0.00\n" % (i + 1, *refc[i]))
        f.write("END\n")
    fr = host.Job(cfg).frames()
    m = masses(names)
    want = com(refc, m, range(NA))
    for f in range(xyz.shape[0]):
        assert np.allclose(com(fr[f].astype(np.float64), m, range(NA)), want, atol=1e-5)

    def variant(a, b):
        text = open(cfg).read()
        assert a in text
        p = os.path.join(str(tmp_path), "v.xml")
        open(p, "w").write(text.replace(a, b))
        return p
    with pytest.raises(host.HostError, match="Fitting routine not understood"):
        host.Job(variant("<type>fittrans</type>", "<type>wiggle</type>"))
    with pytest.raises(host.HostError, match="Must be pre or post"):
        host.Job(variant("<selection>system</selection>\n       <reference>", "<selection>system</selection><order>mid</order>\n       <reference>"))
    with pytest.raises(host.HostError, match="Reference type not understood"):
        host.Job(variant("<type>file</type>", "<type>cloud</type>"))
    with pytest.raises(host.HostError, match="format for alignment reference"):
        host.Job(variant("<format>pdb</format><selection>system</selection></reference>", "<format>dcd</format><selection>system</selection></reference>"))
    with pytest.raises(host.HostError, match="selection not found"):
        host.Job(variant("<type>fittrans</type><selection>system</selection>", "<type>fittrans</type><selection>ghosts</selection>"))


def test_motion_continues_over_clones_and_runs_through_the_hot_path(tmp_path, oracle):
    """clones re-read the frames with new frame numbers, so a linear motion keeps advancing (coordinate_sets.cpp:247 takes
    the GLOBAL frame number); the moved coordinates are what the scatter devices see"""
    from oracle_backend import OracleBackend
    extra = """<motions><motion><type>linear</type><displace>0.5</displace></motion></motions>"""
    fs = "<frameset><file>traj.dcd</file><format>dcd</format><clones>2</clones></frameset>"
    cfg, xyz, names = make_case(tmp_path, NA=16, NF=6, sample_extra=SEL + extra, framesets=fs,
                                scattering="<vectors><type>single</type><single><x>0.7</x><y>0</y><z>0</z></single></vectors>"
                                           "<average><orientation><type>none</type></orientation></average>")
    job = host.Job(cfg)
    fr = job.frames()
    assert fr.shape[0] == 12
    for g in range(12):
        assert np.allclose(fr[g, :, 0] - xyz[g % 6, :, 0], 0.5 * g, atol=1e-5)
        assert np.allclose(fr[g, :, 1:], xyz[g % 6, :, 1:], atol=1e-5)
    be = OracleBackend()
    n, _ = job.run(str(tmp_path / "sig"), backend=be.vtbl)
    assert n == 1
    fqt = np.load(str(tmp_path / "sig" / "fqt.npy"))
    b = job.factors(0.7)
    want = oracle.compute_all_vectors(fr, b, np.array([[0.7, 0.0, 0.0]]), nthreads=2)[0]
    got = fqt[0, :, 0] + 1j * fqt[0, :, 1]
    assert np.max(np.abs(got - want)) < 1e-9 * np.max(np.abs(want))


KINDS = ("linear", "fixed", "oscillation", "randomwalk", "brownian", "localbrownian", "rotationalbrownian")


def test_walkers_pinned_to_reference_build(oracle):
    """tests/golden/ref_smath.npz holds the 4x4 transforms of the reference's own walkers (src/sample/motion_walker.cpp compiled
    where it lies over the uBLAS / Boost.Random shims, oracle/_ref build).  The product's walkers reproduce them bit for bit:
    translation vectors, the sampling discards, the cumulative sums, the redraw loop of localbrownian, the rotation products.
    (The random streams themselves are the shims' Boost-1.4x restatement -- the same one the product uses -- so the streams are
    not pinned, their use is.)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_smath.npz"))
    for i, row in enumerate(g["walk_params"]):
        kw = dict(displace=row[0], frequency=row[1], radius=row[2], seed=int(row[3]), sampling=int(row[4]), direction=tuple(row[5:8]))
        for kind in KINDS:
            want = g[f"walk_{kind}_{i}"]
            got = host.motion_transforms(kind, len(want), **kw)
            assert np.array_equal(got, want), (kind, i)
    if oracle.have_ref_smath():  # live: other parameters
        for kind in KINDS:
            kw = dict(displace=0.9, frequency=0.07, radius=4.0, seed=99, sampling=2, direction=(-1.0, 0.5, 0.25))
            assert np.array_equal(host.motion_transforms(kind, 25, **kw), oracle.ref_motion_transforms(kind, 25, **kw)), kind
    with pytest.raises(host.HostError, match="Motion type not understood"):
        host.motion_transforms("wobble", 3)
    with pytest.raises(host.HostError, match="radius size for local brownian"):
        host.motion_transforms("localbrownian", 3, displace=2.0, radius=1.0)


def _rotfit_case(tmp_path, kind, ref_sel, NA=16, NF=6):
    extra = f"""<alignments><alignment><type>{kind}</type><selection>{ref_sel}</selection><order>post</order>
         <reference><type>frame</type><frame>0</frame><selection>{ref_sel}</selection></reference></alignment></alignments>"""
    return job_frames(tmp_path, extra, NA=NA, NF=NF)


def test_rotational_fit_pinned_to_reference_build(tmp_path, oracle):
    """fitrottrans / fitrot against the REFERENCE's own Fit (src/sample/center_of_mass.cpp:53-157 compiled where it lies into
    oracle/_ref/libparams_ref.so: mass-weighted correlation kernel, LAPACK dgesvd through its Boost.Bindings call -- here the
    OpenBLAS that scipy bundles --, determinant correction, write-back).  The product computes the same rotation with Horn's
    quaternion form; both are narrowed to float by the stager, so they agree to a float ulp.  tests/golden/ref_rotfit.npz holds
    the reference build's output for the case below (tests/golden/make_ref_golden.py); live where the reference is present."""
    import os
    from test_control_plane import REF_DB_NAMES
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_rotfit.npz"))
    for kind in ("fitrottrans", "fitrot"):
        for sel_name, sel in (("system", np.arange(16)), ("front", np.arange(10))):
            d = tmp_path / f"{kind}_{sel_name}"
            d.mkdir()
            job, fr, xyz, names = _rotfit_case(d, kind, sel_name)
            assert np.array_equal(xyz, g["xyz"])
            want = g[f"{kind}_{sel_name}"]
            # prefix selections only: the reference writes cs_redcopy[atom index] back (in bounds and right only then)
            assert np.max(np.abs(fr - want)) < 2e-6, (kind, sel_name, np.max(np.abs(fr - want)))
            if oracle.have_ref_fit():
                for el, rx in REF_DB_NAMES.items():
                    oracle.ref_sample_name_reg(el, rx)
                    oracle.ref_mass_reg(el, MASS[el])
                for f in range(xyz.shape[0]):
                    fit, pc = oracle.ref_fit(str(d / "sample.pdb"), xyz[f], xyz[0], sel, sel)
                    if kind == "fitrot":  # the reference adds the old centre back after the fit (coordinate_sets.cpp:282-286)
                        fit[sel] = fit[sel] + pc
                    assert np.array_equal(fit.astype(np.float32), want[f]), (kind, sel_name, f)
