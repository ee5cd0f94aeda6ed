"""HDF5 signal file ("next" row; reference src/services/file_writer_service.cpp:44-171, 314-484).

libhdf5 / h5py do not exist in this image, so the container code (csrc/host/h5mini.cpp) is pinned in two steps:
  1. the READER must decode a file written by the real library: scipy ships a MATLAB v7.3 file (= HDF5 behind a 512-byte
     user block, written by libhdf5 1.6/1.8 in 2008) whose content scipy's own tests state: testdouble = pi/4 * arange(9);
  2. the WRITER is then checked through that reader (values, extents, max extents, chunk shapes, meta strings), on the
     structural points a libhdf5 reader relies on (signatures, end-of-file address, B-tree node fan-out, sorted links),
     and byte-for-byte against the superblock / heap / symbol-node conventions seen in the real file.
  3. a SECOND, independent reader (tests/h5_independent.py: pure Python, written from the format specification, no code shared
     with h5mini) is checked against the same libhdf5-written file and then reads what h5mini writes -- chunked, extendible
     datasets with one-, two- and three-level chunk B-trees, fresh and resumed files: two implementations of the specification
     have to agree on every byte they dereference.
The chunk B-tree and extendible dataspaces cannot be cross-read by libhdf5 itself here; DESIGN.md says so."""
import os
import struct

import numpy as np
import pytest

from sassena_b200 import host

HERE = os.path.dirname(os.path.abspath(__file__))


def _matlab_file():
    import scipy.io
    p = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(p):
        pytest.skip("scipy's MATLAB v7.3 (HDF5) test file is not installed")
    return p


def test_reader_decodes_a_file_written_by_libhdf5():
    d, lay = host.read_h5(_matlab_file(), with_layout=True)
    assert list(d) == ["testdouble"]
    # scipy/io/matlab/tests/test_mio.py: theta = pi/4 * arange(9), stored by MATLAB as a 9x1 column
    assert d["testdouble"].shape == (9, 1)
    assert np.array_equal(d["testdouble"][:, 0], np.pi / 4 * np.arange(9))
    assert lay["testdouble"] == {"maxdims": None, "chunk": None}


def _signal(n, NF, seed=0):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 3))
    fqt = rng.normal(size=(n, NF)) + 1j * rng.normal(size=(n, NF))
    fq = rng.normal(size=n) + 1j * rng.normal(size=n)
    fq2 = rng.normal(size=n) + 0j
    return q, fqt, fq, fq2


@pytest.mark.parametrize("n,NF,chunksize", [(0, 5, 10000), (1, 1, 10000), (7, 100, 10000), (50, 37, 16), (200, 3, 2), (3, 25, 10)])
def test_signal_file_roundtrip_and_layout(tmp_path, n, NF, chunksize):
    """datasets, extents, unlimited first dimension and chunk shapes follow file_writer_service.cpp:68-168"""
    q, fqt, fq, fq2 = _signal(n, NF)
    p = tmp_path / "signal.h5"
    rows = host.write_signal_h5(p, q, fqt, fq, fq2, chunksize=chunksize, rawconfig="<root/>\n", config="<root/>", database="<database/>")
    assert rows == n
    d, lay = host.read_h5(p, with_layout=True)
    assert sorted(d) == ["fq", "fq0", "fq2", "fqt", "meta/config", "meta/database", "meta/rawconfig", "qvectors"]
    assert d["meta/rawconfig"] == b"<root/>\n" and d["meta/config"] == "<root/>" and d["meta/database"] == "<database/>"
    assert d["qvectors"].shape == (n, 3) and d["fqt"].shape == (n, NF, 2) and d["fq0"].shape == (n, 2)
    s = host.load_signal_h5(p)
    assert np.array_equal(s["qvectors"], q) and np.array_equal(s["fqt"], fqt) and np.array_equal(s["fq"], fq)
    assert np.array_equal(s["fq2"], fq2) and np.array_equal(s["fq0"], fqt[:, 0])
    U = 2**64 - 1
    c2 = NF if 0 < NF < chunksize else chunksize
    assert lay["qvectors"] == {"maxdims": (U, 3), "chunk": (chunksize, 3)}
    assert lay["fqt"] == {"maxdims": (U, NF, 2), "chunk": (max(1, chunksize // c2), c2, 2)}
    for k in ("fq0", "fq", "fq2"):
        assert lay[k] == {"maxdims": (U, 2), "chunk": (chunksize, 2)}
    assert lay["meta/config"] == {"maxdims": None, "chunk": None}


def test_file_structure_matches_libhdf5_conventions(tmp_path):
    """byte-level checks of what a libhdf5 reader dereferences first, compared with the real file's encoding"""
    q, fqt, fq, fq2 = _signal(150, 4)
    p = tmp_path / "s.h5"
    host.write_signal_h5(p, q, fqt, fq, fq2, chunksize=2)  # fqt: 150 x 2 chunks -> a two-level chunk B-tree
    b = open(p, "rb").read()
    real = open(_matlab_file(), "rb").read()[512:]
    assert b[:8] == b"\x89HDF\r\n\x1a\n" == real[:8]
    assert b[8:20] == real[8:20]  # versions 0/0/0, 8-byte offsets and lengths, group K 4 / 16 -- same as the library wrote
    base, free, eof, drv = struct.unpack("<4Q", b[24:56])
    assert (base, free, drv) == (0, 2**64 - 1, 2**64 - 1) and eof == len(b)
    name_off, root_hdr, cache, _, bt, heap = struct.unpack("<QQIIQQ", b[56:96])
    assert (name_off, cache) == (0, 1)
    assert b[bt:bt + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP" and b[root_hdr] == 1
    # root object header: one symbol-table message (type 0x11) pointing at the same B-tree and heap
    nmsg, _, hsize = struct.unpack("<HII", b[root_hdr + 2:root_hdr + 12])
    mtype, msize = struct.unpack("<HH", b[root_hdr + 16:root_hdr + 20])
    assert (nmsg, mtype, msize, hsize) == (1, 0x11, 16, 24)
    assert struct.unpack("<QQ", b[root_hdr + 24:root_hdr + 40]) == (bt, heap)
    # local heap: data segment right after the 32-byte header (as in the real file), free list inside the segment
    seg_size, free_head, seg = struct.unpack("<QQQ", b[heap + 8:heap + 32])
    assert seg == heap + 32 and 8 <= free_head < seg_size
    nxt, fsz = struct.unpack("<QQ", b[seg + free_head:seg + free_head + 16])
    assert nxt == 1 and fsz == seg_size - free_head  # H5HL_FREE_NULL, block runs to the end of the segment
    # group B-tree leaf -> symbol node with the links sorted by name
    level, used = b[bt + 5], struct.unpack("<H", b[bt + 6:bt + 8])[0]
    assert (b[bt + 4], level, used) == (0, 0, 1)
    key0, snod, key1 = struct.unpack("<QQQ", b[bt + 24:bt + 48])
    assert key0 == 0 and b[snod:snod + 4] == b"SNOD" and b[snod + 4] == 1
    nsym = struct.unpack("<H", b[snod + 6:snod + 8])[0]
    names = []
    for i in range(nsym):
        off = struct.unpack("<Q", b[snod + 8 + 40 * i:snod + 16 + 40 * i])[0]
        names.append(b[seg + off:b.index(b"\0", seg + off)].decode())
    assert names == sorted(names) == ["fq", "fq0", "fq2", "fqt", "meta", "qvectors"]
    assert b[seg + key1:b.index(b"\0", seg + key1)].decode() == names[-1]
    # every chunk B-tree node: type 1, at most 64 entries, keys strictly increasing, siblings linked
    pos, nodes = 0, []
    while True:
        pos = b.find(b"TREE\x01", pos)
        if pos < 0:
            break
        nodes.append(pos)
        pos += 4
    assert len(nodes) >= 9  # 5 datasets, fqt alone has 300 chunks = 5 leaves + 1 root
    levels = {}
    for n in nodes:
        lvl, used = b[n + 5], struct.unpack("<H", b[n + 6:n + 8])[0]
        assert used <= 64
        levels.setdefault(lvl, 0)
        levels[lvl] += 1
    assert levels.get(1, 0) >= 1
    d = host.read_h5(p)
    assert np.array_equal(d["fqt"], np.stack([fqt.real, fqt.imag], axis=-1))


def test_resume_appends_and_checks_frames(tmp_path):
    """HDF5WriterService::init on an existing file + get_qvectors (file_writer_service.cpp:21-40,194-211)"""
    q, fqt, fq, fq2 = _signal(9, 12, seed=3)
    p = tmp_path / "signal.h5"
    assert host.write_signal_h5(p, q[:4], fqt[:4], fq[:4], fq2[:4], chunksize=5) == 4
    assert host.write_signal_h5(p, q[4:], fqt[4:], fq[4:], fq2[4:], chunksize=5, resume=True) == 9
    s = host.load_signal_h5(p)
    assert np.array_equal(s["qvectors"], q) and np.array_equal(s["fqt"], fqt) and np.array_equal(s["fq2"], fq2)
    with pytest.raises(host.HostError, match="does not match"):
        host.write_signal_h5(p, q[:1], fqt[:1, :5], fq[:1], fq2[:1], resume=True)  # other number of frames
    (tmp_path / "junk.h5").write_bytes(b"not hdf5 at all" * 10)
    with pytest.raises(host.HostError, match="HDF5 data file"):
        host.write_signal_h5(tmp_path / "junk.h5", q[:1], fqt[:1], fq[:1], fq2[:1], resume=True)
    with pytest.raises(host.HostError, match="Number of frames"):
        host.write_signal_h5(tmp_path / "x.h5", np.zeros((0, 3)), np.zeros((0, 0)), [], [])


def test_job_writes_signal_h5_and_resumes(tmp_path, oracle):
    """scatter.xml -> signal.h5 through the oracle-bound backend; a second run with more q-vectors only computes the new
    ones (sassena.cpp:270-305) and keeps the old rows"""
    from oracle_backend import OracleBackend
    from test_control_plane import ORIENT, make_case
    be = OracleBackend()
    scan = "<vectors><type>scans</type><scans><scan><from>0.5</from><to>1.5</to><points>%d</points><base><x>1</x><y>0</y><z>0</z></base></scan></scans></vectors>"
    cfg, xyz, names = make_case(tmp_path, scattering=scan % 3 + ORIENT)
    sig = tmp_path / "signal.h5"
    job = host.Job(cfg)
    written, _ = job.run(sig, backend=be.vtbl)
    assert written == 3
    d = host.read_h5(sig)
    assert d["meta/rawconfig"].decode() == open(cfg).read() and d["meta/database"] == open(tmp_path / "db.xml").read()
    s = host.load_signal_h5(sig)
    p = job.params()
    for i, q in enumerate(s["qvectors"]):
        fqt, fq, fq2 = oracle.compute_all_vectors(xyz, job.factors(np.linalg.norm(q)), p.init_subvectors(q))
        assert np.allclose(s["fqt"][i], fqt, rtol=1e-11, atol=1e-11 * abs(fqt[0]))
        assert np.isclose(s["fq"][i], fq, rtol=1e-11) and np.isclose(s["fq2"][i], fq2, rtol=1e-11)
        assert s["fq0"][i] == s["fqt"][i][0]
    # same config again: nothing left to compute, file unchanged
    before = open(sig, "rb").read()
    written, _ = host.Job(cfg).run(sig, backend=be.vtbl)
    assert written == 0 and host.read_h5(sig)["qvectors"].shape == (3, 3)
    assert np.array_equal(host.load_signal_h5(sig)["fqt"], s["fqt"]) and len(open(sig, "rb").read()) == len(before)
    # five points: 0.5, 0.75, 1.0, 1.25, 1.5 -> 0.5, 1.0, 1.5 are there already
    cfg5 = tmp_path / "scatter5.xml"
    cfg5.write_text(open(cfg).read().replace("<points>3</points>", "<points>5</points>"))
    job5 = host.Job(str(cfg5))
    written, _ = job5.run(sig, backend=be.vtbl)
    assert written == 2
    s5 = host.load_signal_h5(sig)
    assert s5["qvectors"].shape == (5, 3) and np.array_equal(s5["fqt"][:3], s["fqt"])
    assert sorted(np.round(s5["qvectors"][:, 0], 6)) == [0.5, 0.75, 1.0, 1.25, 1.5]


def test_interrupted_run_is_recovered_from_the_row_journal(tmp_path, oracle):
    """Rows are journaled as the device delivers them (the reference's writer appends as partitions deliver,
    file_writer_service.cpp:314-484).  A run that died after two q-vectors leaves <signal>.h5.d/rows.journal behind and no
    usable signal file; the next run takes those rows over and computes only the rest (sassena.cpp:270-305)."""
    from oracle_backend import OracleBackend
    from test_control_plane import ORIENT, make_case
    be = OracleBackend()
    scan = "<vectors><type>scans</type><scans><scan><from>0.5</from><to>1.5</to><points>5</points><base><x>1</x><y>0</y><z>0</z></base></scan></scans></vectors>"
    cfg, xyz, names = make_case(tmp_path, scattering=scan + ORIENT)
    sig = tmp_path / "signal.h5"
    written, _ = host.Job(cfg).run(sig, backend=be.vtbl)
    assert written == 5 and not os.path.exists(str(sig) + ".d/rows.journal")  # a finished run leaves no journal
    full = host.load_signal_h5(sig)
    NF = full["fqt"].shape[1]
    # the interrupted run: no signal file, a journal with rows 1 and 3 and a torn third record
    os.remove(sig)
    rec = []
    for i in (1, 3):
        rec += list(full["qvectors"][i]) + [full["fq"][i].real, full["fq"][i].imag, full["fq2"][i].real, full["fq2"][i].imag]
        rec += list(full["fqt"][i].view(np.float64))
    blob = np.array([20260117.5, float(NF)] + rec + [1.0, 2.0, 3.0], dtype="<f8").tobytes()
    os.makedirs(str(sig) + ".d", exist_ok=True)
    with open(str(sig) + ".d/rows.journal", "wb") as f:
        f.write(blob)
    written, _ = host.Job(cfg).run(sig, backend=be.vtbl)
    assert written == 3
    got = host.load_signal_h5(sig)
    assert got["qvectors"].shape == (5, 3)
    assert np.array_equal(got["qvectors"][:2], full["qvectors"][[1, 3]]) and np.array_equal(got["fqt"][:2], full["fqt"][[1, 3]])
    order = [int(np.argmin(np.abs(full["qvectors"][:, 0] - q[0]))) for q in got["qvectors"]]
    assert sorted(order) == [0, 1, 2, 3, 4]
    for row, i in enumerate(order):
        assert np.allclose(got["fqt"][row], full["fqt"][i], rtol=1e-13, atol=0) and np.isclose(got["fq2"][row], full["fq2"][i])
    assert not os.path.exists(str(sig) + ".d/rows.journal")


def test_independent_reader_decodes_the_libhdf5_file():
    import h5_independent as hi
    d, lay = hi.read(_matlab_file(), with_layout=True)
    assert list(d) == ["testdouble"] and d["testdouble"].shape == (9, 1)
    assert np.array_equal(d["testdouble"][:, 0], np.pi / 4 * np.arange(9))
    assert lay["testdouble"] == {"maxdims": None, "chunk": None}


@pytest.mark.parametrize("n,NF,chunksize", [(0, 5, 10000), (1, 1, 10000), (7, 100, 10000), (50, 37, 16), (200, 3, 2), (3, 25, 10),
                                            (5000, 4, 2)])
def test_independent_reader_reads_what_h5mini_writes(tmp_path, n, NF, chunksize):
    """the writer's files through the independent reader: same names, values, extents, unlimited dimensions and chunk shapes as
    through h5mini's own reader.  (5000, 4, 2): 10 000 chunks of fqt under a three-level chunk B-tree (fan-out 64)."""
    import h5_independent as hi
    q, fqt, fq, fq2 = _signal(n, NF, seed=3)
    p = tmp_path / "signal.h5"
    host.write_signal_h5(p, q, fqt, fq, fq2, chunksize=chunksize, rawconfig="<root/>\n", config="<root/>", database="<database/>")
    d, lay = hi.read(p, with_layout=True)
    own, own_lay = host.read_h5(p, with_layout=True)
    assert sorted(d) == sorted(own)
    for k in ("qvectors", "fqt", "fq0", "fq", "fq2"):
        assert np.array_equal(d[k], own[k]) and d[k].shape == own[k].shape
        assert lay[k]["maxdims"] == own_lay[k]["maxdims"] and lay[k]["chunk"] == own_lay[k]["chunk"]
    assert np.array_equal(d["qvectors"], q.reshape(n, 3)) and np.array_equal(d["fqt"][..., 0] + 1j * d["fqt"][..., 1], fqt)
    assert np.array_equal(d["fq"][:, 0] + 1j * d["fq"][:, 1], fq) and np.array_equal(d["fq0"][:, 0] + 1j * d["fq0"][:, 1], fqt[:, 0])
    assert d["meta/rawconfig"].tobytes() == b"<root/>\n"
    assert d["meta/config"].tobytes().rstrip(b"\0") == b"<root/>" and d["meta/database"].tobytes().rstrip(b"\0") == b"<database/>"
    f = hi.File(p)
    assert f.eof == os.path.getsize(p) and (f.leaf_k, f.internal_k) == (4, 16)  # group K as libhdf5's defaults


def test_independent_reader_reads_a_resumed_file(tmp_path):
    """rows appended to an existing file (H5Dset_extent + hyperslab writes in the reference, file_writer_service.cpp:314-484)"""
    import h5_independent as hi
    q, fqt, fq, fq2 = _signal(90, 6, seed=5)
    p = tmp_path / "signal.h5"
    assert host.write_signal_h5(p, q[:40], fqt[:40], fq[:40], fq2[:40], chunksize=4) == 40
    assert host.write_signal_h5(p, q[40:], fqt[40:], fq[40:], fq2[40:], chunksize=4, resume=True) == 90
    d = hi.read(p)
    assert d["qvectors"].shape == (90, 3) and np.array_equal(d["qvectors"], q)
    assert np.array_equal(d["fqt"][..., 0] + 1j * d["fqt"][..., 1], fqt)


def test_reader_refuses_implausible_extents_and_cyclic_trees(tmp_path):
    """A corrupt signal file must be refused, not turned into a petabyte allocation or an endless walk: dataspace
    dimensions far beyond the file's size, addresses near 2^64 (which would wrap a naive bounds test) and a chunk B-tree
    whose child pointer leads back to the node itself."""
    rng = np.random.default_rng(3)
    q = rng.standard_normal((6, 3))
    c = lambda *sh: rng.standard_normal(sh) + 1j * rng.standard_normal(sh)  # noqa: E731
    path = tmp_path / "signal.h5"
    host.write_signal_h5(path, q, c(6, 5), c(6), c(6), chunksize=2)
    raw = open(path, "rb").read()
    good = host.read_h5(path)
    assert good["fqt"].shape == (6, 5, 2)
    # (1) blow up the first dimension of fqt's dataspace: find the dims (6, 5, 2) as three little-endian 64-bit words
    import struct
    pat = struct.pack("<3Q", 6, 5, 2)
    at = raw.index(pat)
    for huge in (2**40, 2**62, 2**64 - 1):
        bad = tmp_path / "huge.h5"
        bad.write_bytes(raw[:at] + struct.pack("<Q", huge) + raw[at + 8:])
        with pytest.raises(host.HostError, match="not plausible|beyond"):
            host.read_h5(bad)
    # (2) a chunk B-tree node of level 1 whose children point back at itself
    trees = [i for i in range(0, len(raw) - 8) if raw[i:i + 4] == b"TREE" and raw[i + 4] == 1]
    assert trees
    t = trees[0]
    cyc = bytearray(raw)
    cyc[t + 5] = 1  # level 1: children are nodes
    nd = 4 if struct.unpack_from("<I", raw, t + 24)[0] == 2 * 5 * 2 * 8 else 3  # rank + 1: fqt's tree or fq's / fq2's / fq0's
    keysize = 8 + 8 * nd
    used = struct.unpack_from("<H", raw, t + 6)[0]
    for i in range(used):
        struct.pack_into("<Q", cyc, t + 24 + i * (keysize + 8) + keysize, t)
    bad = tmp_path / "cycle.h5"
    bad.write_bytes(bytes(cyc))
    with pytest.raises(host.HostError, match="levels are inconsistent"):
        host.read_h5(bad)
    # (3) addresses near 2^64
    # (every 8-byte word after the dataspace in turn; the point is only that no mutation crashes or hangs)
    for off in range(at + 24, min(at + 400, len(raw) - 8), 8):
        bad = tmp_path / "addr.h5"
        bad.write_bytes(raw[:off] + struct.pack("<Q", 2**64 - 9) + raw[off + 8:])
        try:
            host.read_h5(bad)
        except host.HostError:
            pass
