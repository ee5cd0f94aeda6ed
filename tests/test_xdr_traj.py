"""XTC / TRR trajectory readers ("next" row; reference src/sample/frames.cpp:592-858).

The reference decodes these files through its vendored xdrfile 1.1.1.  The product's readers (csrc/host/xdr_traj.cpp) are
written from the file formats; here they are pinned against xdrfile itself:
  * committed fixtures written AND decoded by the reference's xdrfile build (tests/golden/make_xdr_golden.py) — bit-exact;
  * where oracle/_ref/libxdrfile_ref.so exists (the build container), a live comparison on fresh random trajectories.
Angstrom conversion follows the reference: (float)(10.0 * (double)nm) (frames.cpp:705-714, data_stager.cpp:111-113)."""
import os
import shutil

import numpy as np
import pytest

from sassena_b200 import host

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
REF_LIB = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libxdrfile_ref.so")

XTC_CASES = ["water", "gas", "tiny", "wide", "coarse", "mixed"]
TRR_CASES = ["trr_single", "trr_vf", "trr_double"]


def to_angstrom(nm):
    return (10.0 * nm.astype(np.float64)).astype(np.float32)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLD, "xdr_golden.npz"))


@pytest.mark.parametrize("name", XTC_CASES + TRR_CASES)
def test_reader_matches_xdrfile_fixture(golden, name):
    ext = "trr" if name.startswith("trr") else "xtc"
    f = host.XdrFile(os.path.join(GOLD, f"xdr_{name}.{ext}"), format=ext)
    ref = golden[name + "_nm"]
    assert (f.number_of_frames, f.number_of_atoms) == ref.shape[:2]
    got, box = f.read(with_box=True)
    assert got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), to_angstrom(ref).view(np.uint32))  # bit-exact
    assert np.array_equal(box, 10.0 * golden[name + "_box_nm"].astype(np.float64))
    # random access and sub-ranges read the same frames
    if f.number_of_frames > 2:
        assert np.array_equal(f.read(2, 1)[0], got[2]) and np.array_equal(f.read(1, 1)[0], got[1])
    f.close()


def test_trim_index_and_errors(golden, tmp_path):
    path = os.path.join(GOLD, "xdr_water.xtc")
    full = host.XdrFile(path).read()
    f = host.XdrFile(path, first=1, last=3, stride=2)  # absolute-index rule of FileFrameset::trim_index: keeps frame 2
    assert f.number_of_frames == 1 and np.array_equal(f.read()[0], full[2])
    with pytest.raises(host.HostError, match="out of bounds"):
        f.read(0, 2)
    with pytest.raises(host.HostError, match="appears not to be a XTC file"):
        host.XdrFile(os.path.join(GOLD, "xdr_trr_single.trr"), format="xtc")
    with pytest.raises(host.HostError, match="appears not to be a TRR file"):
        host.XdrFile(path, format="trr")
    with pytest.raises(host.HostError, match="Unable to open"):
        host.XdrFile(str(tmp_path / "missing.xtc"))
    # a truncated last frame is not indexed (the reference stops at the first frame that does not read back)
    data = open(path, "rb").read()
    cut = tmp_path / "cut.xtc"
    cut.write_bytes(data[:len(data) - 40])
    g = host.XdrFile(str(cut))
    assert g.number_of_frames == full.shape[0] - 1 and np.array_equal(g.read(), full[:-1])


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref (the reference's xdrfile build) only exists in the build container")
@pytest.mark.parametrize("seed,natoms,precision,spread", [(1, 10, 1000.0, 1.0), (2, 999, 1000.0, 4.0), (3, 3000, 500.0, 6.0),
                                                          (4, 64, 10000.0, 30.0), (5, 500, 1.0e5, 200.0), (6, 301, 10.0, 3.0)])
def test_reader_matches_xdrfile_live(tmp_path, seed, natoms, precision, spread):
    """fresh random trajectories (clustered triplets + uniform atoms, negative coordinates, several precisions) written
    and decoded by the reference's xdrfile, read back by the product reader: bit-exact."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_xdr_golden", os.path.join(GOLD, "make_xdr_golden.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    lib = g.load_ref()
    rng = np.random.default_rng(seed)
    nf = 3
    nmol = natoms // 6
    parts = []
    if nmol:
        parts.append(g.water_box(rng, nmol, nf, spread) - 0.3 * spread)
    rest = natoms - 3 * nmol
    parts.append(rng.uniform(-0.5 * spread, spread, (nf, rest, 3)).astype(np.float32))
    xyz = np.concatenate(parts, axis=1)
    xyz = xyz[:, rng.permutation(natoms) if seed % 2 else np.arange(natoms)]
    box = g.boxes(nf, spread)
    p = str(tmp_path / "t.xtc")
    g.write_xtc(lib, p, xyz, box, precision)
    ref, _ = g.read_xtc(lib, p)
    got = host.XdrFile(p).read()
    assert np.array_equal(got.view(np.uint32), to_angstrom(ref).view(np.uint32))
    p = str(tmp_path / "t.trr")
    g.write_trr(lib, p, xyz, box, with_vf=bool(seed % 2))
    ref, _ = g.read_trr(lib, p)
    got = host.XdrFile(p, format="trr").read()
    assert np.array_equal(got.view(np.uint32), to_angstrom(ref).view(np.uint32))


def test_xtc_and_trr_framesets_drive_a_job(tmp_path, golden):
    """sample.framesets with format xtc / trr feed the stager like a DCD (frames.cpp:128-181): selection to
    stager.target, first/stride/clones."""
    from test_control_plane import make_case, SCAN
    ref = to_angstrom(golden["trr_single_nm"])  # 4 frames x 33 atoms
    shutil.copy(os.path.join(GOLD, "xdr_trr_single.trr"), tmp_path / "t.trr")
    cfg, _, _ = make_case(tmp_path, NA=33, NF=2, scattering=SCAN,
                          framesets="<frameset><file>t.trr</file><format>trr</format><first>1</first><clones>2</clones></frameset>")
    job = host.Job(cfg)
    assert job.nframes == 6
    got = job.frames()
    assert np.array_equal(got[:3], ref[1:]) and np.array_equal(got[3:], ref[1:])
    refx = to_angstrom(golden["water_nm"])  # 4 frames x 180 atoms
    shutil.copy(os.path.join(GOLD, "xdr_water.xtc"), tmp_path / "w.xtc")
    cfg, _, _ = make_case(tmp_path, NA=180, NF=2, scattering=SCAN,
                          framesets="<frameset><file>w.xtc</file><format>xtc</format><stride>2</stride></frameset>")
    job = host.Job(cfg)
    assert job.nframes == 2 and np.array_equal(job.frames(), refx[::2])
    cfg, _, _ = make_case(tmp_path, NA=12, NF=2, scattering=SCAN,
                          framesets="<frameset><file>w.xtc</file><format>xtc</format></frameset>")
    with pytest.raises(host.HostError, match="Atom number mismatch"):
        host.Job(cfg)
