"""Mutation fuzzing of the file readers of the host layer (trajectories: DCD, XTC, TRR; signal files: HDF5 subset; the text
inputs of a job: scatter.xml, db.xml, PDB structure, ndx selections), meant to
run against the ASan + UBSan build (tools/asan_host.sh builds it; see tools/fuzz_readers.sh).  A mutated file must either read
or raise host.HostError -- anything else (another exception, a sanitizer report, a crash, a hang) is a finding.

Usage: python tools/fuzz_readers.py [iterations per file] [seed]
"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sassena_b200 import host  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def mutate(data: bytes, rng) -> bytes:
    b = bytearray(data)
    kind = rng.integers(0, 4)
    if kind == 0 and len(b) > 1:  # truncate
        return bytes(b[:rng.integers(0, len(b))])
    if kind == 1:  # flip a few bytes
        for _ in range(rng.integers(1, 9)):
            b[rng.integers(0, len(b))] = rng.integers(0, 256)
        return bytes(b)
    if kind == 2:  # overwrite an aligned 32-bit word with an extreme value (sizes, counts, magic numbers live there)
        for _ in range(rng.integers(1, 4)):
            o = 4 * rng.integers(0, max(1, len(b) // 4))
            b[o:o + 4] = [b"\xff\xff\xff\xff", b"\x7f\xff\xff\xff", b"\x00\x00\x00\x00", b"\x80\x00\x00\x00",
                          b"\xff\xff\xff\x7f", b"\x00\x00\x00\x80"][rng.integers(0, 6)]
        return bytes(b)
    # splice: drop or duplicate a block
    if len(b) > 16:
        o = rng.integers(0, len(b) - 8)
        n = rng.integers(1, min(4096, len(b) - o))
        return bytes(b[:o] + b[o + n:]) if rng.integers(0, 2) else bytes(b[:o + n] + b[o:])
    return bytes(b)


def read_traj(path, fmt):
    f = host.DCDFile(path) if fmt == "dcd" else host.XdrFile(path, format=fmt)
    try:
        n = f.number_of_frames
        if n:
            f.read(0, min(n, 64))
    finally:
        f.close()


TOKENS = [b"<", b">", b"</", b"/>", b"&", b"&amp;", b"<!--", b"-->", b"<![CDATA[", b"]]>", b"\"", b"'", b"\0", b"-1", b"1e999", b"nan",
          b"99999999999999999999", b"<scan>", b"</sample>", b"<xi:include href=\"scatter.xml\"/>", b"\n", b" " * 100]


def mutate_text(data: bytes, rng) -> bytes:
    b = bytearray(data)
    kind = rng.integers(0, 5)
    if kind == 0:
        return bytes(b[:rng.integers(0, len(b))])
    if kind == 1:
        for _ in range(rng.integers(1, 6)):
            b[rng.integers(0, len(b))] = rng.integers(0, 256)
        return bytes(b)
    if kind == 2:  # insert a token that means something to an XML / number / column parser
        o = rng.integers(0, len(b))
        return bytes(b[:o] + TOKENS[rng.integers(0, len(TOKENS))] + b[o:])
    if kind == 3:  # delete a span
        o = rng.integers(0, len(b))
        return bytes(b[:o] + b[o + rng.integers(1, 40):])
    o = rng.integers(0, len(b))  # duplicate a span
    n = rng.integers(1, 200)
    return bytes(b[:o + n] + b[o:])


def fuzz_job(tmp, iters, rng):
    """scatter.xml, db.xml, the PDB structure and an ndx selection file of a small job, one mutated at a time"""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    from test_control_plane import make_case, SCAN, ORIENT  # noqa: E402
    d = os.path.join(tmp, "job")
    os.makedirs(d)
    extra = """<selections>
      <selection><type>index</type><name>picked</name><index>3</index><index>5</index></selection>
      <selection><type>range</type><name>tail</name><from>20</from><to>23</to></selection>
      <selection><type>lexical</type><name>carbons</name><expression>carbon</expression></selection>
      <selection><type>file</type><name>flagged</name><file>sel.pdb</file><format>pdb</format></selection>
      <selection><type>file</type><file>groups.ndx</file><format>ndx</format><expression>grp.*</expression></selection>
    </selections>
    <motions><motion><type>linear</type><displace>0.5</displace><selection>picked</selection></motion></motions>
    <alignments><alignment><type>center</type><selection>carbons</selection></alignment></alignments>"""
    cfg, _, _ = make_case(d, scattering=SCAN + ORIENT, sample_extra=extra,
                          background="<background><factor>0.1</factor></background>")
    open(os.path.join(d, "sel.pdb"), "w").write(open(os.path.join(d, "sample.pdb")).read())
    open(os.path.join(d, "groups.ndx"), "w").write("[ grpA ]\n1 2 3\n4\n[ other ]\n7 8\n[ grpB ]\n10 11\n")
    findings = 0
    for target in ("scatter.xml", "db.xml", "sample.pdb", "groups.ndx"):
        path = os.path.join(d, target)
        data = open(path, "rb").read()
        ok = err = 0
        t0 = time.time()
        for i in range(iters):
            with open(path, "wb") as fh:
                fh.write(mutate_text(data, rng))
            try:
                job = host.Job(cfg)
                try:
                    q = job.qvectors()
                    if len(q):
                        job.factors(float(np.linalg.norm(q[0])))
                    job.frames()
                finally:
                    job.close()
                ok += 1
            except host.HostError:
                err += 1
            except Exception as e:  # noqa: BLE001
                findings += 1
                keep = os.path.join(tmp, f"finding_job_{findings}_{target}")
                os.replace(path, keep)
                print(f"FINDING {type(e).__name__}: {e} -> {keep}", flush=True)
        with open(path, "wb") as fh:
            fh.write(data)
        print(f"{target:24s} {iters} mutations: {ok} read, {err} rejected, {time.time() - t0:.1f} s", flush=True)
    return findings


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    tmp = tempfile.mkdtemp(prefix="fuzz_")
    seeds = []
    for name in sorted(os.listdir(GOLD)):
        if name.endswith(".xtc") or name.endswith(".trr"):
            seeds.append((os.path.join(GOLD, name), name[-3:]))
    dcd = os.path.join(tmp, "seed.dcd")
    host.write_dcd(dcd, rng.standard_normal((7, 33, 3)).astype(np.float32))
    seeds.append((dcd, "dcd"))
    h5 = os.path.join(tmp, "seed.h5")
    q = rng.standard_normal((5, 3))
    c = lambda *sh: rng.standard_normal(sh) + 1j * rng.standard_normal(sh)  # noqa: E731
    host.write_signal_h5(h5, q, c(5, 9), c(5), c(5), chunksize=2)
    seeds.append((h5, "h5"))
    findings = 0
    for path, fmt in seeds:
        data = open(path, "rb").read()
        ok = err = 0
        t0 = time.time()
        for i in range(iters):
            m = os.path.join(tmp, f"m.{fmt}")
            with open(m, "wb") as fh:
                fh.write(mutate(data, rng))
            try:
                if fmt == "h5":
                    host.read_h5(m)
                else:
                    read_traj(m, fmt)
                ok += 1
            except host.HostError:
                err += 1
            except Exception as e:  # noqa: BLE001
                findings += 1
                keep = os.path.join(tmp, f"finding_{findings}.{fmt}")
                os.replace(m, keep)
                print(f"FINDING {type(e).__name__}: {e} -> {keep}", flush=True)
        print(f"{os.path.basename(path):24s} {iters} mutations: {ok} read, {err} rejected, {time.time() - t0:.1f} s", flush=True)
    findings += fuzz_job(tmp, iters, rng)
    if not findings:  # the mutated files of a finding stay for inspection
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    print("findings:", findings)
    return 1 if findings else 0


if __name__ == "__main__":
    sys.exit(main())
