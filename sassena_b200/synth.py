"""Synthetic workloads of SURVEY.md 8(d) and the CPU twin of the device trajectory generator.

Host-side numpy only (inputs, not compute).  The random walk uses an integer hash (splitmix64) and exactly
rounded float32 operations so that `trajectory()` here and `sgpu_synth_trajectory` (csrc/kernels/amplitude.cu)
produce bit-identical coordinates: large benchmark inputs are generated on the GPU while the oracle checks a
sub-sample generated here.
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return z ^ (z >> np.uint64(31))


def _hash(seed, atom, t, c):
    with np.errstate(over="ignore"):
        inner = _splitmix64(np.uint64(seed) ^ (atom * np.uint64(3) + np.uint64(c)))
        return _splitmix64(inner + np.uint64(t) * np.uint64(0xD1342543DE82EF95))


def step_scale(sigma: float) -> np.float32:
    """Irwin-Hall(4) of 16-bit uniforms has variance 65536^2/3 -> scale so a step has std sigma."""
    return np.float32(sigma * np.sqrt(3.0) / 65536.0)


def trajectory(NF, NA, box, sigma, seed, atoms=None, offset=0.0, layout=0):
    """float32 coordinates, layout 0: [NF][n][3], 1: [n][NF][3], for `atoms` (default all NA)."""
    atoms = np.arange(NA, dtype=np.uint64) if atoms is None else np.asarray(atoms, dtype=np.uint64)
    n = len(atoms)
    out = np.empty((NF, n, 3), dtype=np.float32)
    box_scale = np.float32(np.float32(box) / np.float32(16777216.0))
    off = np.float32(offset)
    sc = step_scale(sigma)
    pos = np.empty((n, 3), dtype=np.float32)
    with np.errstate(over="ignore"):
        for c in range(3):
            h = _hash(seed, atoms, 0xFFFFFFFF, c)
            pos[:, c] = (h >> np.uint64(40)).astype(np.float32) * box_scale + off
        out[0] = pos
        for t in range(1, NF):
            for c in range(3):
                h = _hash(seed, atoms, t, c)
                isum = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)) +
                        ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + ((h >> np.uint64(48)) & np.uint64(0xFFFF)))
                step = (isum.astype(np.int64) - 131070).astype(np.float32) * sc
                pos[:, c] = pos[:, c] + step
            out[t] = pos
    return out if layout == 0 else np.ascontiguousarray(out.transpose(1, 0, 2))


def factors(NA):
    """b_j in {-3.74 (H), 6.65 (C), 5.80 (O)} by j mod 3."""
    return np.array([-3.74, 6.65, 5.80])[np.arange(NA) % 3]


def unit_vectors(n, seed):
    """normalised N(0,1)^3 rows; passed explicitly (= vectors.type=file) because the Boost sphere stream is
    version dependent."""
    v = np.random.default_rng(seed).normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def qlengths(frm, to, points):
    """|q| list of one scan along (1,0,0) with the reference's float-rounded fractions (parameters.cpp:1151)."""
    if points == 1:
        return np.array([(frm + to) / 2])
    out = [frm]
    for j in range(1, points - 1):
        out.append(frm + float(np.float32(np.float32(j * 1.0 / (points - 1)) ** np.float32(1.0))) * (to - frm))
    out.append(to)
    return np.array(out[:points])


# the five BASELINE.json configurations
CONFIGS = {
    "C1": dict(kind="all", NA=1000, NF=100, box=30.0, sigma=0.1, seed=1, q=(0.2, 2.0, 10), NM=100, vseed=2),
    "C2": dict(kind="self", NA=30000, NF=10000, box=70.0, sigma=0.05, seed=3, q=(0.1, 2.0, 20), NM=200, vseed=4),
    "C3": dict(kind="all", NA=100000, NF=10000, box=100.0, sigma=0.05, seed=5, q=(0.1, 5.0, 50), NM=500, vseed=6),
    "C4": dict(kind="mpsphere", NA=1000000, NF=1000, box=220.0, sigma=0.05, seed=7, q=(0.01, 0.5, 200), L=20,
               offset=-110.0),
    "C5": dict(kind="self", NA=500000, NF=50000, box=170.0, sigma=0.05, seed=8, q=(0.1, 2.0, 20), NM=200, vseed=9),
}
