"""Golden vectors produced by the three generators of the REFERENCE's own src/control/parameters.cpp (:930-1189): orientation
vectors (sphere / cylinder boost_uniform_on_sphere, cylinder raster_linear, file), multipole moments (sphere / cylinder), and
q-vector scans (powf fractions, 1-3 nested scans), compiled where it lies into oracle/_ref/libparams_ref.so over the shims
(oracle/Makefile, entry points oracle/ref_params_wrap.cpp).  Boost.Random is absent: the uniform_on_sphere cases run over the
Boost-1.4x restatement in oracle/shim/boost/random, so they pin the generators' USE of the stream (dimension, draws per vector,
z = 0 for the cylinder), not the stream.  Run in the build container (needs /root/reference); writes
tests/golden/ref_params.npz, which travels with the repo.

    python tests/golden/make_ref_params_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as o  # noqa: E402
import test_reference_params as t  # noqa: E402

subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
out = {}
for seed in t.SEEDS:
    out[f"sphere_{seed}"] = o.ref_orientation_vectors("sphere", resolution=t.NV, seed=seed)
    out[f"cylinder_{seed}"] = o.ref_orientation_vectors("cylinder", resolution=t.NV, seed=seed)
for r in t.RASTER:
    out[f"raster_{r}"] = o.ref_orientation_vectors("cylinder", "raster_linear", r)
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "qqq.txt")
    np.savetxt(p, t.FILE_VECTORS)
    out["vectors_file"] = o.ref_orientation_vectors("file", filepath=p)
    p = os.path.join(d, "mm.txt")
    np.savetxt(p, t.FILE_MOMENTS, fmt="%d")
    out["moments_file"] = o.ref_multipole_moments("sphere", type="file", filepath=p)
for L in t.MOMENT_RES:
    out[f"moments_sphere_{L}"] = o.ref_multipole_moments("sphere", L)
    out[f"moments_cylinder_{L}"] = o.ref_multipole_moments("cylinder", L)
for i, s in enumerate(t.SCANS):
    out[f"scan_{i}"] = o.ref_scan_vectors(s)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_params.npz"), **out)
print("wrote tests/golden/ref_params.npz:", len(out), "arrays")
