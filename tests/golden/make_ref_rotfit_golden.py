"""Golden vectors of the REFERENCE's own rotational fit: oracle/_ref/libparams_ref.so holds src/sample/center_of_mass.cpp compiled
where it lies (make -C oracle ref; uBLAS / Boost.Bindings shims, dgesvd from the OpenBLAS scipy bundles).  Run in the build
container (needs /root/reference); writes tests/golden/ref_rotfit.npz, which travels with the repo and pins the product's
fitrottrans / fitrot alignments on machines without the reference (tests/test_coordinate_sets.py).

    python tests/golden/make_ref_rotfit_golden.py
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
from test_control_plane import REF_DB_NAMES, make_case  # noqa: E402
from test_coordinate_sets import MASS  # noqa: E402

subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
assert o.have_ref_fit()
for el, rx in REF_DB_NAMES.items():
    o.ref_sample_name_reg(el, rx)
    o.ref_mass_reg(el, MASS[el])
out = {}
with tempfile.TemporaryDirectory() as tmp:
    cfg, xyz, names = make_case(tmp, NA=16, NF=6)  # the case of test_rotational_fit_pinned_to_reference_build
    xyz = xyz.astype(np.float64)
    out["xyz"] = xyz
    for kind in ("fitrottrans", "fitrot"):
        for sel_name, sel in (("system", np.arange(16)), ("front", np.arange(10))):
            frames = []
            for f in range(xyz.shape[0]):
                fit, pc = o.ref_fit(os.path.join(tmp, "sample.pdb"), xyz[f], xyz[0], sel, sel)
                if kind == "fitrot":  # coordinate_sets.cpp:282-286: the old centre of mass is added back after the fit
                    fit[sel] = fit[sel] + pc
                frames.append(fit.astype(np.float32))  # the stager's narrowing (data_stager.cpp:111-113)
            out[f"{kind}_{sel_name}"] = np.array(frames)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_rotfit.npz"), **out)
print("wrote tests/golden/ref_rotfit.npz:", {k: v.shape for k, v in out.items()})
