"""Control plane ("next" row): scatter.xml + db.xml + PDB + DCD -> the hot path.

CPU tests check the native parser/database/selection/frameset logic (csrc/host/control.cpp) against restatements of the
reference's formulas written here (database.cpp:391-528, scatter_factors.cpp:56-78) and run the whole flow through the
oracle-bound backend table; the GPU test runs the same job on the CUDA path and compares with the oracle."""
import math
import os

import numpy as np
import pytest

from sassena_b200 import host, synth

ELEMENTS = ["hydrogen", "carbon", "oxygen", "nitrogen"]
# atom names cycle through these PDB names
PDB_NAMES = ["H1", "CA", "OW", "N", "HB2", "C", "O", "CB"]
NAME2EL = {"H1": "hydrogen", "HB2": "hydrogen", "CA": "carbon", "C": "carbon", "CB": "carbon", "OW": "oxygen",
           "O": "oxygen", "N": "nitrogen"}

DB_XML = """<?xml version="1.0"?>
<!-- small test database, same schema as the reference's db.xml -->
<database>
  <names><pdb>
    <element><name>hydrogen</name><param>^H.*</param></element>
    <element><name>carbon</name><param>^C.*</param></element>
    <element><name>oxygen</name><param>^O.*</param></element>
    <element><name>nitrogen</name><param>^N.*</param></element>
  </pdb></names>
  <masses>
    <element><name>hydrogen</name><param>1.008</param></element>
    <element><name>carbon</name><param>12.011</param></element>
    <element><name>oxygen</name><param>15.999</param></element>
    <element><name>nitrogen</name><param>14.007</param></element>
  </masses>
  <sizes>
    <element><name>hydrogen</name><type>1</type><param>1.07</param></element>
    <element><name>carbon</name><type>1</type><param>1.58</param></element>
    <element><name>oxygen</name><type>2</type><param>1.3</param></element>
    <element><name>nitrogen</name><type>0</type><param>11.5</param></element>
  </sizes>
  <exclusionfactors>
    <element><name>hydrogen</name><type>2</type><param>1.0</param></element>
    <element><name>carbon</name><type>2</type><param>0.9</param></element>
    <element><name>oxygen</name><type>1</type><param>1.1</param></element>
    <element><name>nitrogen</name><type>0</type><param>3.0</param></element>
  </exclusionfactors>
  <scatterfactors>
    <element><name>hydrogen</name><type>0</type><param>-3.7406</param></element>
    <element><name>carbon</name><type>2</type>
      <param>2.31</param><param>20.8439</param><param>1.02</param><param>10.2075</param>
      <param>1.5886</param><param>0.5687</param><param>0.865</param><param>51.6512</param><param>0.2156</param></element>
    <element><name>oxygen</name><type>1</type>
      <param>7.6579</param><param>2.2458</param><param>2.2266</param><param>0</param><param>0</param>
      <param>2</param><param>2</param><param>4</param><param>0</param><param>0</param>
      <param>1</param><param>2</param><param>2</param><param>1</param><param>1</param></element>
    <element><name>nitrogen</name><type>0</type><param>9.36</param></element>
  </scatterfactors>
</database>
"""


def pdb_line(i, name, beta=0.0, segid="PROT"):
    # columns: 1-6 record, 13-16 name, 61-66 beta, 73-76 segid
    nm = name if len(name) == 4 else " " + name.ljust(3)
    return "ATOM  %5d %4s ALA A%4d    %8.3f%8.3f%8.3f%6.2f%6.2f      %-4s" % (i + 1, nm, 1, 0.0, 0.0, 0.0, 1.0, beta, segid)


def make_case(tmp, NA=24, NF=20, scattering="", sample_extra="", stager="", framesets=None, background=""):
    tmp = str(tmp)
    names = [PDB_NAMES[i % len(PDB_NAMES)] for i in range(NA)]
    with open(os.path.join(tmp, "sample.pdb"), "w") as f:
        f.write("REMARK test structure\n")
        for i, n in enumerate(names):
            f.write(pdb_line(i, n, beta=1.0 if i % 3 == 0 else 0.0, segid="SOLV" if i >= NA // 2 else "PROT") + "\n")
        f.write("END\n")
    with open(os.path.join(tmp, "db.xml"), "w") as f:
        f.write(DB_XML)
    xyz = synth.trajectory(NF, NA, 15.0, 0.3, 11, offset=-7.5)
    host.write_dcd(os.path.join(tmp, "traj.dcd"), xyz)
    if framesets is None:
        framesets = "<frameset><file>traj.dcd</file><format>dcd</format></frameset>"
    cfg = f"""<?xml version="1.0" encoding="UTF-8"?>
<root>
  <!-- comment before the sample -->
  <sample>
    <structure><file>sample.pdb</file><format>pdb</format></structure>
    <framesets>{framesets}</framesets>
    {sample_extra}
  </sample>
  {stager}
  <scattering>
    {background}
    {scattering}
  </scattering>
  <database><file>db.xml</file></database>
</root>
"""
    path = os.path.join(tmp, "scatter.xml")
    with open(path, "w") as f:
        f.write(cfg)
    return path, xyz, names


# ---- python restatement of the database evaluation (float32 powf as in the reference) ----
import ctypes as _C
import ctypes.util as _Cu

_libm = _C.CDLL(_Cu.find_library("m"))
_libm.powf.restype = _C.c_float
_libm.powf.argtypes = [_C.c_float, _C.c_float]


def powf(a, b):  # the C library's float pow the reference calls (database.cpp:400-506)
    return float(_libm.powf(a, b))


def ref_volume(el):
    if el == "hydrogen":
        return (4.0 / 3.0) * math.pi * powf(1.07, 3)
    if el == "carbon":
        return (4.0 / 3.0) * math.pi * powf(1.58, 3)
    if el == "oxygen":
        # `using namespace std` + float argument: the reference's sqrt is the float overload and the product of the
        # two floats is a float product (database.cpp:404)
        return float(np.sqrt(np.float32(powf(math.pi, 3))) * np.float32(powf(1.3, 3)))
    return 11.5


def ref_excl(el, effv, q):
    if el in ("hydrogen", "carbon"):
        v0 = 1.0 if el == "hydrogen" else 0.9
        return effv * math.exp(-1.0 * powf(effv, 2.0 / 3.0) * powf(q, 2) / (4 * math.pi)) * v0
    if el == "oxygen":
        return effv * 1.1
    return 3.0


def ref_sfactor(el, q):
    if el == "hydrogen":
        return -3.7406
    if el == "nitrogen":
        return 9.36
    if el == "carbon":
        v = [2.31, 20.8439, 1.02, 10.2075, 1.5886, 0.5687, 0.865, 51.6512, 0.2156]
        return sum(v[2 * i] * math.exp(-v[2 * i + 1] * q * q) for i in range(4)) + v[8]
    v = [7.6579, 2.2458, 2.2266, 0, 0, 2, 2, 4, 0, 0, 1, 2, 2, 1, 1]
    den = 0.0
    for j in range(5):
        j2, j3 = 5 + j, 10 + j
        if v[j] != 0.0:
            c = powf(2.0 * v[j] / v[j3], v[j3] + 0.5)
            c /= math.sqrt(math.factorial(int(2.0 * v[j3])))
            den += v[j2] * powf(c * powf(q, v[j3] - 1.0) * math.exp(-1.0 * v[j] * q / v[j3]), 2)
    return den / (4.0 * math.pi)


# ---- the same tables for the reference's own Database (oracle/_ref build of src/control/database.cpp) ----
REF_DB_ELEMENTS = ["hydrogen", "carbon", "oxygen", "nitrogen"]
REF_DB_TABLES = {
    "sizes": {"hydrogen": (1, [1.07]), "carbon": (1, [1.58]), "oxygen": (2, [1.3]), "nitrogen": (0, [11.5])},
    "exclusionfactors": {"hydrogen": (2, [1.0]), "carbon": (2, [0.9]), "oxygen": (1, [1.1]), "nitrogen": (0, [3.0])},
    "scatterfactors": {"hydrogen": (0, [-3.7406]),
                       "carbon": (2, [2.31, 20.8439, 1.02, 10.2075, 1.5886, 0.5687, 0.865, 51.6512, 0.2156]),
                       "oxygen": (1, [7.6579, 2.2458, 2.2266, 0, 0, 2, 2, 4, 0, 0, 1, 2, 2, 1, 1]), "nitrogen": (0, [9.36])},
}
REF_DB_NAMES = {"hydrogen": "^H.*", "carbon": "^C.*", "oxygen": "^O.*", "nitrogen": "^N.*"}
REF_DB_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_database.npz")


def register_reference_database(o):
    """DB_XML's tables registered through the reference's own reg() methods (its XML reader needs libxml2)"""
    for table, rows in REF_DB_TABLES.items():
        for i, el in enumerate(REF_DB_ELEMENTS):
            o.ref_db_reg(table, i, rows[el][1], rows[el][0])
    for el, rx in REF_DB_NAMES.items():
        o.ref_db_name_reg(el, rx)


SCAN = """<vectors><type>scans</type><scans>
  <scan><from>0.4</from><to>1.6</to><points>3</points><base><x>1</x><y>0</y><z>0</z></base></scan>
</scans></vectors>"""


def test_config_selections_framesets_and_scans(tmp_path):
    sel_pdb = tmp_path / "sel.pdb"
    ndx = tmp_path / "groups.ndx"
    sample_extra = """<selections>
      <selection><type>index</type><name>picked</name><index>3</index><index>5</index><index>8</index></selection>
      <selection><type>range</type><name>tail</name><from>20</from><to>23</to></selection>
      <selection><type>lexical</type><name>carbons</name><expression>carbon</expression></selection>
      <selection><type>file</type><name>flagged</name><file>sel.pdb</file><format>pdb</format></selection>
      <selection><type>file</type><name>solvent</name><file>sel.pdb</file><format>pdb</format>
                 <selector>segid</selector><expression>SOLV</expression></selection>
      <selection><type>file</type><file>groups.ndx</file><format>ndx</format><expression>grp.*</expression></selection>
      <selection><type>index</type><name>system</name><index>1</index></selection>
    </selections>"""
    framesets = """<first>2</first><stride>2</stride>
      <frameset><file>traj.dcd</file><format>dcd</format><last>12</last></frameset>
      <frameset><file>traj.dcd</file><format>dcd</format><first>0</first><stride>5</stride><clones>2</clones></frameset>"""
    cfg, xyz, names = make_case(tmp_path, scattering=SCAN, sample_extra=sample_extra, framesets=framesets,
                                stager="<stager><target>carbons</target></stager>")
    # selection files live next to the config (get_filepath, parameters.cpp:41-54)
    sel_pdb.write_text((tmp_path / "sample.pdb").read_text())
    ndx.write_text("[ grpA ]\n1 2 3\n4\n[ other ]\n7 8\n[ grpB ]\n10 11\n")
    job = host.Job(cfg)
    NA = len(names)
    assert job.natoms == NA
    assert list(job.selection("picked")) == [3, 5, 8]
    assert list(job.selection("tail")) == [20, 21, 22, 23]
    carbons = [i for i, n in enumerate(names) if NAME2EL[n] == "carbon"]
    assert list(job.selection("carbons")) == carbons
    assert list(job.selection("flagged")) == [i for i in range(NA) if i % 3 == 0]
    assert list(job.selection("solvent")) == list(range(NA // 2, NA))
    assert list(job.selection("grpA")) == [0, 1, 2, 3] and list(job.selection("grpB")) == [9, 10]
    with pytest.raises(host.HostError):
        job.selection("other")
    # the reserved name is moved aside and "system" is every atom (sample.cpp:88-101)
    assert list(job.selection("system")) == list(range(NA))
    assert list(job.selection("system_RENAMED_BY_SASSENA")) == [1]
    # framesets: defaults first=2 stride=2 (absolute i % stride), frameset 1 last=12 -> 2,4,..,12; frameset 2 overrides
    # first=0 stride=5 -> 0,5,10,15 twice (clones)
    idx = [2, 4, 6, 8, 10, 12] + [0, 5, 10, 15] * 2
    assert job.nframes == len(idx) and job.ntarget == len(carbons)
    assert np.array_equal(job.frames(), xyz[idx][:, carbons])
    # scans
    q = host.create_from_scans([{"base": (1, 0, 0), "from": 0.4, "to": 1.6, "points": 3}])
    assert job.nqvectors == 3 and np.array_equal(job.qvectors(), q)


def test_selections_equal_reference_readers(tmp_path, oracle):
    """structure and selection readers against the REFERENCE's own atoms.cpp / atomselection_reader.cpp / atomselection.cpp
    (oracle/_ref/libparams_ref.so; live only where the reference is present -- the readers take files, so there is no fixture
    beyond the expectations spelled out in test_config_selections_framesets_and_scans)"""
    if not oracle.have_ref_params():
        pytest.skip("oracle/_ref/libparams_ref.so not built (no /root/reference on this machine)")
    sample_extra = """<selections>
      <selection><type>range</type><name>tail</name><from>20</from><to>23</to></selection>
      <selection><type>lexical</type><name>carbons</name><expression>carbon</expression></selection>
      <selection><type>lexical</type><name>heavy</name><expression>(carbon|oxygen|nitrogen)</expression></selection>
      <selection><type>file</type><name>flagged</name><file>sel.pdb</file><format>pdb</format></selection>
      <selection><type>file</type><name>b0</name><file>sel.pdb</file><format>pdb</format><selector>beta</selector>
                 <expression>0\\.00</expression></selection>
      <selection><type>file</type><name>solvent</name><file>sel.pdb</file><format>pdb</format>
                 <selector>segid</selector><expression>SOLV</expression></selection>
      <selection><type>file</type><name>anyseg</name><file>sel.pdb</file><format>pdb</format>
                 <selector>segid</selector><expression>(SOLV|PROT)</expression></selection>
      <selection><type>file</type><file>groups.ndx</file><format>ndx</format><expression>grp.*</expression></selection>
    </selections>"""
    cfg, xyz, names = make_case(tmp_path, scattering=SCAN, sample_extra=sample_extra)
    pdb = str(tmp_path / "sel.pdb")
    ndx = str(tmp_path / "groups.ndx")
    (tmp_path / "sel.pdb").write_text((tmp_path / "sample.pdb").read_text())
    (tmp_path / "groups.ndx").write_text("[ grpA ]\n1 2 3\n4\n[ other ]\n7 8\n[grpB]\n10 11\n\n[ grpC ] trailing\n24\n")
    job = host.Job(cfg)
    for el, rx in REF_DB_NAMES.items():
        oracle.ref_sample_name_reg(el, rx)
    labels = oracle.ref_atoms_labels(str(tmp_path / "sample.pdb"))  # Atoms::add: PDB names -> database labels
    assert labels == [NAME2EL[n] for n in names] and job.natoms == len(labels)
    # lexical: the product selects the INDICES of the atoms whose label matches (the reference pushes the element ID)
    assert list(job.selection("carbons")) == [i for i, l in enumerate(labels) if l == "carbon"]
    assert list(job.selection("heavy")) == [i for i, l in enumerate(labels) if l != "hydrogen"]
    assert np.array_equal(job.selection("tail"), oracle.ref_select_range(20, 23))
    assert np.array_equal(job.selection("flagged"), oracle.ref_select_pdb(pdb, "beta", "1|1\\.0|1\\.00"))
    assert np.array_equal(job.selection("b0"), oracle.ref_select_pdb(pdb, "beta", "0\\.00"))
    assert np.array_equal(job.selection("solvent"), oracle.ref_select_pdb(pdb, "segid", "SOLV"))
    assert np.array_equal(job.selection("anyseg"), oracle.ref_select_pdb(pdb, "segid", "(SOLV|PROT)"))
    # "[grpB]" without blanks is named "grpB]" by the reference (it keeps the closing bracket, atomselection_reader.cpp:51-54)
    for grp in ("grpA", "grpB]", "grpC"):
        assert np.array_equal(job.selection(grp), oracle.ref_select_ndx(ndx, "name", "grp.*", grp)), grp
    assert list(job.selection("grpB]")) == [9, 10] and list(job.selection("grpC")) == [23]
    assert oracle.ref_select_ndx(ndx, "name", "grp.*", "other") is None and oracle.ref_select_ndx(ndx, "name", "grp.*", "grpB") is None
    for absent in ("other", "grpB"):
        with pytest.raises(host.HostError):
            job.selection(absent)


def test_pdb_frameset(tmp_path):
    """PDBFrameset (frames.cpp:442-577): frames end at END/ENDMDL lines, an unterminated last frame counts, the
    ENDMDL+END tail does not add a frame; first/stride/clones apply as for DCD; coordinates come from columns 31-54."""
    cfg, xyz, names = make_case(tmp_path, NA=12, NF=6, scattering=SCAN,
                                framesets="<frameset><file>traj.pdb</file><format>pdb</format><first>1</first>"
                                          "<stride>2</stride><clones>2</clones></frameset>")
    with open(tmp_path / "traj.pdb", "w") as f:
        f.write("CRYST1   30.000   30.000   30.000  90.00  90.00  90.00 P 1           1\n")
        for fr in range(6):
            f.write("MODEL %8d\n" % (fr + 1))
            for i, n in enumerate(names):
                x, y, z = xyz[fr, i]
                nm = n if len(n) == 4 else " " + n.ljust(3)
                f.write("ATOM  %5d %4s ALA A%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n" % (i + 1, nm, 1, x, y, z, 1.0, 0.0))
            f.write("ENDMDL\n" if fr < 5 else "")  # last frame left unterminated
    job = host.Job(cfg)
    # frames 1, 3, 5 pass (i >= 1, i % 2 == 0 never true for odd i -> absolute index rule keeps i = 2, 4)
    keep = [i for i in range(6) if i >= 1 and i % 2 == 0]
    assert job.nframes == 2 * len(keep)
    expect = np.round(xyz[keep].astype(np.float64), 3).astype(np.float32)
    got = job.frames()
    assert np.allclose(got[:len(keep)], expect, atol=6e-4) and np.array_equal(got[:len(keep)], got[len(keep):])
    # a terminated tail ("ENDMDL" + "END") does not create an empty frame; a wrong atom count raises
    with open(tmp_path / "traj.pdb", "a") as f:
        f.write("ENDMDL\nEND\n")
    assert host.Job(cfg).nframes == 2 * len(keep)
    with open(tmp_path / "traj.pdb", "a") as f:
        f.write("ATOM      1  CA  ALA A   1       1.000   2.000   3.000  1.00  0.00\nEND\n")
    with pytest.raises(host.HostError, match="Atom number mismatch"):
        host.Job(cfg)


def test_scatter_factors_match_reference_formulas(tmp_path):
    bg = """<background><factor>0.0334</factor><kappas>
      <kappa><selection>solvent</selection><value>1.5</value></kappa></kappas></background>"""
    extra = """<selections><selection><type>range</type><name>solvent</name><from>12</from><to>23</to></selection>
      <selection><type>range</type><name>half</name><from>6</from><to>17</to></selection></selections>"""
    cfg, xyz, names = make_case(tmp_path, scattering=SCAN, background=bg, sample_extra=extra,
                                stager="<stager><target>half</target></stager>")
    job = host.Job(cfg)
    target = list(range(6, 18))
    for ql in (0.0, 0.7, 2.3):
        exp = []
        for i in target:
            el = NAME2EL[names[i]]
            kappa = 1.5 if i >= 12 else 1.0
            exp.append(ref_sfactor(el, ql) - 0.0334 * ref_excl(el, kappa * ref_volume(el), ql))
        got = job.factors(ql)
        assert np.allclose(got, exp, rtol=1e-14, atol=0), (ql, got, exp)
    # no background: plain scattering lengths
    cfg2, _, names = make_case(tmp_path, scattering=SCAN)
    j2 = host.Job(cfg2)
    assert np.allclose(j2.factors(1.1), [ref_sfactor(NAME2EL[n], 1.1) for n in names], rtol=1e-14)


def test_database_restatement_equals_reference_database(oracle):
    """the formulas restated above (and with them the powf / float-sqrt roundings they claim) against values evaluated by the
    reference's own database.cpp (tests/golden/ref_database.npz, tests/golden/make_ref_database_golden.py): bit for bit; live
    where oracle/_ref is built"""
    g = np.load(REF_DB_GOLD)
    assert list(g["elements"]) == REF_DB_ELEMENTS
    bg = float(g["background"])
    for i, el in enumerate(REF_DB_ELEMENTS):
        assert ref_volume(el) == g["volume"][i], el
        for c, ql in enumerate(g["q"]):
            assert ref_sfactor(el, float(ql)) == g["sfactor"][i, c], (el, ql)
            for k, kappa in enumerate(g["kappa"]):
                e = ref_excl(el, float(kappa) * ref_volume(el), float(ql))
                assert e == g["exclusion"][i, k, c], (el, ql, kappa)
                assert ref_sfactor(el, float(ql)) - bg * e == g["effective"][i, k, c], (el, ql, kappa)
    assert [NAME2EL[n] for n in g["names"]] == list(g["resolved"])
    if oracle.have_ref_smath():
        register_reference_database(oracle)
        for i, el in enumerate(REF_DB_ELEMENTS):
            assert oracle.ref_db_volume(i) == g["volume"][i]
            assert oracle.ref_db_effective(i, 0.9, 1.25, 0.05) == \
                ref_sfactor(el, 0.9) - 0.05 * ref_excl(el, 1.25 * ref_volume(el), 0.9)
        assert all(oracle.ref_db_name_get(n) == NAME2EL[n] for n in NAME2EL)


def test_product_scatter_factors_equal_reference_database(tmp_path):
    """csrc/host/control.cpp (Database + ScatterFactors::update through sass_job_factors) against the reference's own
    database.cpp values: bit for bit, for every atom of the target selection, with and without background / kappas"""
    g = np.load(REF_DB_GOLD)
    bgxml = """<background><factor>0.0334</factor><kappas>
      <kappa><selection>solvent</selection><value>1.5</value></kappa></kappas></background>"""
    extra = """<selections><selection><type>range</type><name>solvent</name><from>12</from><to>23</to></selection></selections>"""
    cfg, _, names = make_case(tmp_path, scattering=SCAN, background=bgxml, sample_extra=extra)
    job = host.Job(cfg)
    el_index = {e: i for i, e in enumerate(REF_DB_ELEMENTS)}
    for c, ql in enumerate(g["q"]):
        got = job.factors(float(ql))
        exp = [g["effective"][el_index[NAME2EL[n]], 1 if i >= 12 else 0, c] for i, n in enumerate(names)]
        assert np.array_equal(got, exp), (ql, got, exp)


def test_config_errors(tmp_path):
    cfg, _, _ = make_case(tmp_path, scattering=SCAN)
    text = open(cfg).read()

    def variant(repl, by):
        p = str(tmp_path / "bad.xml")
        assert repl in text
        open(p, "w").write(text.replace(repl, by))
        return p

    with pytest.raises(host.HostError, match="cannot open"):
        host.Job(str(tmp_path / "nope.xml"))
    with pytest.raises(host.HostError, match="XML parse error"):
        host.Job(variant("</root>", "</toor>"))
    with pytest.raises(host.HostError, match="Selection type not understood"):
        host.Job(variant("</framesets>", "</framesets><selections><selection><type>magic</type></selection></selections>"))
    with pytest.raises(host.HostError, match="Motion type not understood"):
        host.Job(variant("</framesets>", "</framesets><motions><motion><type>wobble</type></motion></motions>"))
    with pytest.raises(host.HostError, match="stager.target"):
        host.Job(variant("<scattering>", "<stager><target>nobody</target></stager><scattering>"))
    with pytest.raises(host.HostError, match="No q vectors"):
        host.Job(variant("<points>3</points>", "<points>0</points>"))
    with pytest.raises(host.HostError, match="not supported"):
        host.Job(variant("<frameset><file>traj.dcd</file><format>dcd</format>", "<frameset><file>traj.dcd</file><format>crd</format>"))
    with pytest.raises(host.HostError, match="appears not to be a XTC file"):
        host.Job(variant("<frameset><file>traj.dcd</file><format>dcd</format>", "<frameset><file>traj.dcd</file><format>xtc</format>"))
    with pytest.raises(host.HostError, match="obsolete"):
        host.Job(variant("<scattering>", "<scattering><target>system</target>"))
    # unknown and ambiguous atom names (database.cpp:308-340)
    pdb = str(tmp_path / "sample.pdb")
    good = open(pdb).read()
    open(pdb, "w").write(good.replace(" N  ", " XE ", 1))
    with pytest.raises(host.HostError, match="not recognized"):
        host.Job(cfg)
    open(pdb, "w").write(good)
    db = str(tmp_path / "db.xml")
    open(db, "w").write(DB_XML.replace("^N.*", "^[NC].*"))
    with pytest.raises(host.HostError, match="matches"):
        host.Job(cfg)
    # atom count mismatch between structure and trajectory
    open(db, "w").write(DB_XML)
    open(pdb, "w").write(good.replace("END\n", pdb_line(99, "CA") + "\nEND\n"))
    with pytest.raises(host.HostError, match="Atom number mismatch"):
        host.Job(cfg)


def test_xinclude(tmp_path):
    """the reference runs libxml2's XInclude processing on every configuration document (xml_interface.cpp:109-116): an
    <xi:include href=...> stands for the root element of the referenced document (relative to the including file, nested
    includes relative to theirs), parse="text" for its text, <xi:fallback> for a missing one"""
    cfg, _, _ = make_case(tmp_path, scattering=SCAN)
    text = open(cfg).read()
    ref_q = host.Job(cfg).qvectors()
    a, b = text.index("<scattering>"), text.index("</scattering>") + len("</scattering>")
    os.makedirs(tmp_path / "parts" / "deep", exist_ok=True)
    scat = text[a:b]
    va, vb = scat.index("<vectors>"), scat.index("</vectors>") + len("</vectors>")
    open(tmp_path / "parts" / "deep" / "vectors.xml", "w").write('<?xml version="1.0"?>\n' + scat[va:vb])
    open(tmp_path / "parts" / "scattering.xml", "w").write(
        '<?xml version="1.0"?>\n' + scat[:va] + '<xi:include xmlns:xi="http://www.w3.org/2001/XInclude" href="deep/vectors.xml"/>'
        + scat[vb:])
    inc = '<xi:include xmlns:xi="http://www.w3.org/2001/XInclude" href="parts/scattering.xml"/>'
    p = str(tmp_path / "with_include.xml")
    open(p, "w").write(text[:a] + inc + text[b:])
    assert np.array_equal(host.Job(p).qvectors(), ref_q)
    # fallback content of a missing document; text inclusion
    fb = ('<xi:include xmlns:xi="http://www.w3.org/2001/XInclude" href="parts/none.xml"><xi:fallback>' + scat
          + '</xi:fallback></xi:include>')
    open(p, "w").write(text[:a] + fb + text[b:])
    assert np.array_equal(host.Job(p).qvectors(), ref_q)
    open(tmp_path / "parts" / "three.txt", "w").write("3")
    assert "<points>3</points>" in text
    open(p, "w").write(text.replace("<points>3</points>", '<points><xi:include xmlns:xi="http://www.w3.org/2001/XInclude" '
                                    'href="parts/three.txt" parse="text"/></points>'))
    assert np.array_equal(host.Job(p).qvectors(), ref_q)
    with pytest.raises(host.HostError, match="XInclude: cannot open"):
        open(p, "w").write(text[:a] + inc.replace("scattering.xml", "absent.xml") + text[b:])
        host.Job(p)
    with pytest.raises(host.HostError, match="xpointer"):
        open(p, "w").write(text[:a] + inc.replace("href=", 'xpointer="x" href=') + text[b:])
        host.Job(p)
    open(tmp_path / "parts" / "loop.xml", "w").write('<a><xi:include xmlns:xi="http://www.w3.org/2001/XInclude" href="loop.xml"/></a>')
    with pytest.raises(host.HostError, match="nested deeper"):
        open(p, "w").write(text[:a] + inc.replace("scattering.xml", "loop.xml") + text[b:])
        host.Job(p)


ORIENT = """<average><orientation><type>vectors</type>
  <vectors><type>sphere</type><algorithm>boost_uniform_on_sphere</algorithm><resolution>7</resolution><seed>5</seed></vectors>
</orientation></average>"""


def _expected(oracle, job, xyz_t, dsp="autocorrelate"):
    p = job.params()
    out = []
    for q in job.qvectors():
        b = job.factors(np.linalg.norm(q))
        out.append(oracle.compute_all_vectors(xyz_t, b, p.init_subvectors(q), dsp=dsp))
    return out


def test_job_run_oracle_backend_writes_signal(tmp_path, oracle):
    from oracle_backend import OracleBackend
    extra = "<selections><selection><type>lexical</type><name>heavy</name><expression>carbon|oxygen|nitrogen</expression></selection></selections>"
    bg = "<background><factor>0.02</factor></background>"
    cfg, xyz, names = make_case(tmp_path, scattering=SCAN + ORIENT + "<signal><fq2>false</fq2></signal>", sample_extra=extra,
                                background=bg, stager="<stager><target>heavy</target></stager>")
    job = host.Job(cfg)
    heavy = [i for i, n in enumerate(names) if NAME2EL[n] != "hydrogen"]
    out = tmp_path / "signal"
    be = OracleBackend()
    written, report = job.run(out, backend=be.vtbl)
    assert written == 3 and "target=%d" % len(heavy) in report
    sig = host.load_signal(out)
    assert "fq2" not in sig and sig["fqt"].shape == (3, xyz.shape[0])
    assert np.array_equal(sig["qvectors"], job.qvectors())
    exp = _expected(oracle, job, xyz[:, heavy])
    for i, (fqt, fq, fq2) in enumerate(exp):
        assert np.allclose(sig["fqt"][i], fqt, rtol=1e-12, atol=1e-12 * abs(fqt[0]))
        assert np.isclose(sig["fq"][i], fq, rtol=1e-12) and sig["fq0"][i] == sig["fqt"][i][0]


@pytest.mark.gpu
def test_job_run_gpu_matches_oracle(tmp_path, oracle, gpu_ctx):
    from util import TOL, rel_err
    bg = "<background><factor>0.0334</factor></background>"
    cfg, xyz, names = make_case(tmp_path, NA=200, NF=64, scattering=SCAN + ORIENT, background=bg)
    job = host.Job(cfg)
    out = tmp_path / "signal"
    written, _ = job.run(out, ctx=gpu_ctx)
    assert written == 3
    sig = host.load_signal(out)
    for i, (fqt, fq, fq2) in enumerate(_expected(oracle, job, xyz)):
        assert rel_err(sig["fqt"][i], fqt) < TOL
        assert abs(sig["fq"][i] - fq) <= TOL * abs(fq) and abs(sig["fq2"][i] - fq2) <= TOL * abs(fq2)


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path, oracle):
    import subprocess
    import sys
    dsp = "<dsp><type>square</type></dsp>"
    cfg, xyz, names = make_case(tmp_path, NA=64, NF=32, scattering=SCAN + ORIENT + dsp)
    out = tmp_path / "sig"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "sassena_b200.cli", "--config", cfg, "--signal", str(out)], cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    from util import TOL, rel_err
    sig = host.load_signal(out)
    job = host.Job(cfg)
    for i, (fqt, fq, fq2) in enumerate(_expected(oracle, job, xyz, dsp="square")):
        assert rel_err(sig["fqt"][i], fqt) < TOL
    # default output: scattering.signal.file = signal.h5 next to the configuration, in the reference's HDF5 layout
    r = subprocess.run([sys.executable, "-m", "sassena_b200.cli", "--config", cfg], cwd=root, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    assert job.signal_file == str(tmp_path / "signal.h5")
    h5 = host.load_signal_h5(job.signal_file)
    assert np.array_equal(h5["qvectors"], sig["qvectors"]) and np.array_equal(h5["fqt"], sig["fqt"])
    assert np.array_equal(h5["fq0"], sig["fqt"][:, 0]) and np.array_equal(h5["fq2"], sig["fq2"])


@pytest.mark.parametrize("kind", ["all", "self"])
def test_stager_dump_writes_the_staged_coordinates(tmp_path, oracle, kind):
    """stager.dump (data_stager.cpp:91-94,131-165,233-238,352-391): the staged coordinates of stager.target as a DCD file --
    frame-major for the coherent / multipole devices, and NA "frames" of NF "atoms" (the atom-major staging layout) for
    the self device"""
    from oracle_backend import OracleBackend
    cfg, xyz, names = make_case(
        tmp_path, NA=12, NF=7,
        sample_extra="<selections><selection><type>range</type><name>part</name><from>2</from><to>9</to></selection></selections>",
        stager="<stager><target>part</target><dump>true</dump><file>staged.dcd</file><format>dcd</format></stager>",
        scattering=f"<type>{kind}</type><vectors><type>single</type><single><x>0.5</x><y>0</y><z>0</z></single></vectors>"
                   "<average><orientation><type>none</type></orientation></average>")
    job = host.Job(cfg)
    fr = job.frames()
    assert fr.shape == (7, 8, 3)
    n, _ = job.run(str(tmp_path / "sig"), backend=OracleBackend().vtbl)
    assert n == 1
    got = host.DCDFile(str(tmp_path / "staged.dcd")).read()
    if kind == "all":
        assert got.shape == fr.shape and np.array_equal(got, fr)
    else:
        assert got.shape == (8, 7, 3) and np.array_equal(got, fr.transpose(1, 0, 2))
    # unknown format: the reference's error
    bad = str(tmp_path / "bad.xml")
    open(bad, "w").write(open(cfg).read().replace("<format>dcd</format></stager>", "<format>xyz</format></stager>"))
    with pytest.raises(host.HostError, match="Format for coordinate dumping not known"):
        host.Job(bad).run(str(tmp_path / "sig2"), backend=OracleBackend().vtbl)


def test_command_line_overwrite_options(tmp_path):
    """Params::options / overwrite_options (parameters.cpp:795-875): the eight overwrite options are applied after the
    configuration file has been read.  (In the reference the three FILE options only change `.file`, which nothing reads
    after the configuration has been parsed; here they take effect, resolved like the configuration's own elements.)"""
    extra = """<selections>
      <selection><type>lexical</type><name>carbons</name><expression>carbon</expression></selection>
      <selection><type>range</type><name>head</name><from>0</from><to>4</to></selection>
    </selections>"""
    cfg, xyz, names = make_case(tmp_path, scattering=SCAN, sample_extra=extra, stager="<stager><target>carbons</target></stager>")
    carbons = [i for i, n in enumerate(names) if NAME2EL[n] == "carbon"]
    base = host.Job(cfg)
    assert base.ntarget == len(carbons) and base.signal_file == str(tmp_path / "signal.h5")
    assert base.option("stager.dump") == "false" and base.option("stager.target") == "carbons"
    # stager.target, stager.dump / file / format, scattering.signal.file
    job = host.Job(cfg, {"stager.target": "head", "stager.dump": True, "stager.file": "out/d.dcd", "stager.format": "dcd",
                         "scattering.signal.file": "run2.h5", "limits.computation.threads": 4})
    assert job.ntarget == 5 and np.array_equal(job.frames(), xyz[:, :5])
    assert job.signal_file == str(tmp_path / "run2.h5")
    assert job.option("stager.dump") == "true" and job.option("stager.file") == str(tmp_path / "out" / "d.dcd")
    assert job.option("stager.target") == "head" and job.option("limits.computation.threads") is None
    # sample.structure.file: a second structure with other atom names -> other elements behind the same selection name
    other = tmp_path / "other.pdb"
    lines = (tmp_path / "sample.pdb").read_text().splitlines()
    other.write_text("\n".join(pdb_line(i, "CA") if ln.startswith("ATOM") else ln
                               for i, ln in enumerate(l for l in lines)) + "\n")
    job2 = host.Job(cfg, {"sample.structure.file": "other.pdb"})
    assert job2.ntarget == len(names)  # every atom is a carbon now
    with pytest.raises(host.HostError, match="structure format not supported"):
        host.Job(cfg, {"sample.structure.format": "gro"})
    with pytest.raises(host.HostError, match="unrecognised option '--stager.mode'"):
        host.Job(cfg, {"stager.mode": "atoms"})
    with pytest.raises(host.HostError, match="is invalid"):
        host.Job(cfg, {"stager.dump": "maybe"})
    with pytest.raises(host.HostError, match="is invalid"):
        host.Job(cfg, {"limits.computation.threads": "many"})
    # the command line of the executable carries them
    from sassena_b200 import cli
    got = {}

    class _Stop(Exception):
        pass

    def fake_job(config, overwrites=None):
        got.update(config=config, overwrites=overwrites)
        raise _Stop

    orig = host.Job
    host.Job = fake_job
    try:
        with pytest.raises(_Stop):
            cli.main(["--config", cfg, "--stager.target", "head", "--scattering.signal.file", "x.h5", "--stager.dump", "1"])
    finally:
        host.Job = orig
    assert got == {"config": cfg, "overwrites": {"stager.target": "head", "scattering.signal.file": "x.h5", "stager.dump": "1"}}


@pytest.mark.parametrize("mode", ["frames", "atoms"])
def test_stage_only_is_the_reference_s_stage_executable(tmp_path, oracle, mode):
    """s_stage (src/main/s_stage.cpp:205-232): stage the trajectory of stager.target by stager.mode and, with stager.dump,
    write the post-processed coordinates -- no scattering calculation, no signal file"""
    from oracle_backend import OracleBackend
    cfg, xyz, names = make_case(
        tmp_path, NA=10, NF=6,
        sample_extra="<selections><selection><type>range</type><name>part</name><from>1</from><to>7</to></selection></selections>",
        stager=f"<stager><target>part</target><mode>{mode}</mode><file>staged.dcd</file></stager>", scattering=SCAN)
    job = host.Job(cfg, {"stager.dump": True})
    fr = job.frames()
    nbytes, report = job.stage(backend=OracleBackend().vtbl)
    assert nbytes == fr.size * 4 and f"stager.mode={mode}" in report and "staged.dcd" in report
    got = host.DCDFile(str(tmp_path / "staged.dcd")).read()
    if mode == "frames":
        assert np.array_equal(got, fr)
    else:
        assert np.array_equal(got, fr.transpose(1, 0, 2))
    assert not os.path.exists(tmp_path / "signal.h5")
    # without the dump nothing is written; an unknown mode is the reference's error
    os.remove(tmp_path / "staged.dcd")
    host.Job(cfg).stage(backend=OracleBackend().vtbl)
    assert not os.path.exists(tmp_path / "staged.dcd")
    bad = str(tmp_path / "bad.xml")
    open(bad, "w").write(open(cfg).read().replace(f"<mode>{mode}</mode>", "<mode>both</mode>"))
    with pytest.raises(host.HostError, match="Staging mode not understood stager.mode=both"):
        host.Job(bad).stage(backend=OracleBackend().vtbl)
