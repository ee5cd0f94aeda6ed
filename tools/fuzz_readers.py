"""Mutation fuzzing of the file readers of the host layer (trajectories: DCD, XTC, TRR; signal files: HDF5 subset), meant to
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
    print("findings:", findings)
    return 1 if findings else 0


if __name__ == "__main__":
    sys.exit(main())
