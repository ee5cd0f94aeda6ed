"""Trajectory readers against the REFERENCE's own frames.cpp (DCDFrameset, PDBFrameset, XTCFrameset, TRRFrameset: header
parsing, generate_index, FileFrameset::trim_index, read_frame, over its vendored xdrfile), compiled where it lies into
oracle/_ref/libparams_ref.so (oracle/ref_frames_wrap.cpp).  The readers take files, so these run live where the reference is
present; on other machines the committed XTC / TRR fixtures (tests/test_xdr_traj.py, written by the reference's xdrfile) and the
reference-written DCD (tests/golden/ref_writer.dcd) hold the pins.  The reference's Frame holds coor2_t doubles (DCD: the file's
floats; XTC / TRR: 10.0 * nm evaluated in double; PDB: strtod of the columns) which its stagers narrow to coor_t float
(data_stager.cpp:111-113); the product reads straight into the float32 staging layout, so the comparison is made after that
narrowing and is exact."""
import os

import numpy as np
import pytest

from sassena_b200 import host, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRIMS = [dict(), dict(first=3, last=9, last_set=True, stride=2), dict(first=1, stride=3), dict(first=0, last=0, last_set=True),
         dict(first=5, last=400, last_set=True, stride=4)]


@pytest.fixture
def ref(oracle):
    if not oracle.have_ref_params():
        pytest.skip("oracle/_ref/libparams_ref.so not built (no /root/reference on this machine)")
    return oracle


def _kw(t):  # the product's readers take last=None for "not set"
    k = dict(first=t.get("first", 0), stride=t.get("stride", 1))
    if t.get("last_set"):
        k["last"] = t["last"]
    return k


def test_dcd_reader_equals_reference(ref, tmp_path):
    xyz = synth.trajectory(13, 9, 25.0, 0.4, 3, offset=-7.0)
    path = str(tmp_path / "traj.dcd")
    host.write_dcd(path, xyz)
    for p in (path, os.path.join(GOLD, "ref_writer.dcd")):
        for t in TRIMS:
            r = ref.ref_frames_read("dcd", p, **t)
            d = host.DCDFile(p, **_kw(t))
            assert d.number_of_frames == len(r), (p, t)
            if len(r):
                assert np.array_equal(d.read(), r.astype(np.float32)), (p, t)
    assert np.array_equal(ref.ref_frames_read("dcd", path), xyz.astype(np.float64))


@pytest.mark.parametrize("name,ext", [("water", "xtc"), ("tiny", "xtc"), ("wide", "xtc"), ("mixed", "xtc"), ("gas", "xtc"),
                                      ("coarse", "xtc"), ("trr_single", "trr"), ("trr_double", "trr"), ("trr_vf", "trr")])
def test_xtc_trr_readers_equal_reference(ref, name, ext):
    p = os.path.join(GOLD, f"xdr_{name}.{ext}")
    for t in TRIMS[:3]:
        r = ref.ref_frames_read(ext, p, **t)
        f = host.XdrFile(p, format=ext, **_kw(t))
        assert f.number_of_frames == len(r), (name, t)
        if len(r):
            assert np.array_equal(f.read(), r.astype(np.float32)), (name, t)


def test_pdb_frameset_equals_reference(ref, tmp_path):
    """multi-model PDB through a whole job (the product reads PDB framesets inside the control plane)"""
    import test_control_plane as t
    cfg, xyz, names = t.make_case(tmp_path, NF=7, scattering=t.SCAN,
                                  framesets="<frameset><file>traj.pdb</file><format>pdb</format><first>1</first></frameset>")
    with open(tmp_path / "traj.pdb", "w") as f:
        f.write("CRYST1   30.000   30.000   30.000  90.00  90.00  90.00 P 1           1\n")
        for fr in range(7):
            f.write("MODEL %8d\n" % (fr + 1))
            for i, n in enumerate(names):
                x, y, z = xyz[fr, i]
                nm = n if len(n) == 4 else " " + n.ljust(3)
                f.write("ATOM  %5d %4s ALA A%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n" % (i + 1, nm, 1, x, y, z, 1.0, 0.0))
            f.write("ENDMDL\n" if fr < 6 else "")  # last frame left unterminated
    r = ref.ref_frames_read("pdb", str(tmp_path / "traj.pdb"), first=1)
    job = host.Job(cfg)
    assert job.nframes == len(r) == 6
    assert np.array_equal(job.frames(), r.astype(np.float32))


def test_coordinate_set_conversions_equal_reference(ref):
    """the oracle's cart_to_spherical / cart_to_cylindrical (the inputs of the multipole devices and the checker of the device-side
    conversion kernels) against the reference's own CartesianCoordinateSet(frame, selection) -> translate -> Spherical /
    CylindricalCoordinateSet (coordinate_set.cpp:278-315) after the stager's narrowing to float: bit for bit, including atoms at
    the origin, on the axes and in every quadrant"""
    xyz = synth.trajectory(1, 64, 30.0, 0.5, 3, offset=-15.0)[0]
    xyz[0] = 0
    xyz[1, :2] = 0
    xyz[2, 0] = 0
    xyz[3, 1] = 0
    xyz[4] = [-1, 0, 0]
    xyz[5] = [0, -2, 0]
    xyz[6] = [0, 0, -3]
    sel = np.array([0, 1, 2, 3, 4, 5, 6, 7, 11, 30, 49, 63])
    assert np.array_equal(ref.ref_coordinate_set(xyz, sel), xyz[sel].astype(np.float64))
    t = np.array([1.5, -2.25, 0.125])
    assert np.array_equal(ref.ref_coordinate_set(xyz, sel, trans=t), xyz[sel].astype(np.float64) + t)
    assert np.array_equal(ref.ref_coordinate_set(xyz, sel, repr="spherical").astype(np.float32), ref.cart_to_spherical(xyz[sel]))
    assert np.array_equal(ref.ref_coordinate_set(xyz, repr="spherical").astype(np.float32), ref.cart_to_spherical(xyz))
    for axis in ((0, 0, 1), (0.3, -0.2, 1.0), (1, 1, 0), (0, 1, 0), (-1, 0, 0)):
        assert np.array_equal(ref.ref_coordinate_set(xyz, sel, repr="cylindrical", axis=axis).astype(np.float32),
                              ref.cart_to_cylindrical(xyz[sel], axis)), axis
