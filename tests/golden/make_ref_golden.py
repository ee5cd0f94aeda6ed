"""Golden vectors produced by the REFERENCE's own code: oracle/_ref/libsmath_ref.so is the reference's src/math/smath.cpp,
src/math/coor3d.cpp, src/decomposition/assignment.cpp, src/decomposition/decomposition_plan.cpp,
src/stager/coordinate_writer.cpp and src/sample/motion_walker.cpp compiled where they lie (make -C oracle ref, shims in oracle/shim).
Run in the build container (needs /root/reference); writes tests/golden/ref_smath.npz, which travels with the repo and
pins the oracle's restatements on machines without the reference (tests/test_oracle.py::test_oracle_pinned_to_reference_build).

    python tests/golden/make_ref_golden.py
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
rng = np.random.default_rng(20261017)
out = {}
# correlation / square: timelines of several lengths (odd, even, prime, power of two), complex128
for i, NF in enumerate((1, 2, 3, 8, 17, 100, 257)):
    x = rng.normal(size=NF) + 1j * rng.normal(size=NF)
    out[f"corr_{i}_x"] = x
    out[f"corr_{i}_direct"] = o.ref_auto_correlate_direct(x)
    out[f"corr_{i}_direct_vec"] = o.ref_auto_correlate_direct(x, vector_overload=True)
    out[f"corr_{i}_fftw"] = o.ref_auto_correlate_fftw(x)
    out[f"corr_{i}_square"] = o.ref_square_elements(x)
# coordinate conversions: random points, points on the axes / in the coordinate planes / at the origin
pts = np.concatenate([rng.normal(size=(200, 3)) * 30.0,
                      np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 2, 0], [0, -2, 0], [0, 0, 3], [0, 0, -3], [1, 1, 0], [-1, 1, 0],
                                [-1, -1, 0], [1, -1, 0], [0, 1, 1], [0, -1, -1], [1e-30, 0, 1], [5, 0, -1e-20]], dtype=np.float64)])
pts = pts.astype(np.float32).astype(np.float64)  # what a frame holds
out["pts"] = pts
out["sph"] = o.ref_cart_to_spherical(pts)
axes = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0], [1, 1, 0], [2, -1, 3], [0, 0, -2], [-1, -1, -1], [1e-3, 0, 1]], dtype=np.float64)
out["axes"] = axes
out["bases"] = np.array([o.ref_vector_base(a) for a in axes])
out["cyl"] = np.array([o.ref_cart_to_cylindrical(pts, a) for a in axes])
# assignments: (NN, NAF) grid, every rank
rows = []
for NN in (1, 2, 3, 4, 7, 8, 16):
    for NAF in (1, 2, 5, 8, 9, 23, 100, 1000):
        for rank in range(NN):
            for mod in (0, 1):
                off, size, mx, idx = o.ref_assignment(bool(mod), NN, rank, NAF)
                first = int(idx[0]) if size else -1
                last = int(idx[-1]) if size else -1
                rows.append((mod, NN, rank, NAF, off, size, mx, first, last, int(idx.sum())))
out["assignments"] = np.array(rows, dtype=np.int64)
# decomposition: penalties on a grid, and the automatic / manual plans wherever one exists (the oracle's own search screens
# the inputs: the reference reports "no plan" with a bare `throw;` that would terminate this script)
pen = []
for NN in (1, 2, 3, 4, 8, 16, 64):
    for NQ in (1, 2, 5, 20, 50):
        for NAF in (1, 7, 100, 10000):
            for NNpP in sorted({1, 2, 3, NN // 2 if NN > 1 else 1, NN}):
                if NNpP <= NN:
                    pen.append((NN, NQ, NAF, NNpP, o.ref_decomposition_penalty(NN, NQ, NAF, NNpP)))
out["penalties"] = np.array(pen, dtype=np.int64)
plans = []
for NN in (1, 2, 3, 4, 8, 16, 64):
    for NQ in (1, 2, 5, 20, 50):
        for NAF in (1, 7, 100, 10000):
            for maxbytes in (10 ** 12, 12 * 100000 * 20):
                el = 12 * 100000
                rc, part, psize, _ = o.decomposition_plan(NN, NQ, NAF, el, maxbytes, 0.0)
                if rc != 0:
                    continue
                rp, rps, rpen, col = o.ref_decomposition_plan(NN, NQ, NAF, el, maxbytes, True, 1, 0.0)
                plans.append((NN, NQ, NAF, el, maxbytes, 1, 1, rp, rps, rpen, int((col * (np.arange(NN) + 1)).sum())))
            for manual in (1, 2, 5):
                rp, rps, rpen, col = o.ref_decomposition_plan(NN, NQ, NAF, 12, 10 ** 12, False, manual, 0.0)
                plans.append((NN, NQ, NAF, 12, 10 ** 12, 0, manual, rp, rps, rpen, int((col * (np.arange(NN) + 1)).sum())))
out["plans"] = np.array(plans, dtype=np.int64)
# a DCD file out of the reference's own DCDCoordinateWriter (src/stager/coordinate_writer.cpp), written in two pieces
from sassena_b200 import synth  # noqa: E402
dcd_xyz = synth.trajectory(7, 13, 20.0, 0.3, 5)
out["dcd_xyz"] = dcd_xyz
o.ref_dcd_write(os.path.join(ROOT, "tests", "golden", "ref_writer.dcd"), dcd_xyz, split=3)
# motion walkers: the reference's own src/sample/motion_walker.cpp over the uBLAS / Boost.Random shims (the random streams are
# the shims' Boost-1.4x restatement; what is pinned is how the walkers consume and accumulate them)
WALK = [dict(displace=0.37, frequency=0.013, radius=1.5, seed=7, sampling=3, direction=(1.0, 2.0, -2.0)),
        dict(displace=2.0, frequency=0.25, radius=20.0, seed=12345, sampling=1, direction=(0.0, 0.0, 1.0))]
for kind in ("linear", "fixed", "oscillation", "randomwalk", "brownian", "localbrownian", "rotationalbrownian"):
    for i, kw in enumerate(WALK):
        out[f"walk_{kind}_{i}"] = o.ref_motion_transforms(kind, 60, **kw)
out["walk_params"] = np.array([[k["displace"], k["frequency"], k["radius"], k["seed"], k["sampling"], *k["direction"]] for k in WALK])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_smath.npz"), **out)
print("wrote tests/golden/ref_smath.npz:", len(out), "arrays,", len(rows), "assignment rows,", len(pen), "penalties,", len(plans), "plans")
