"""Golden values evaluated by the REFERENCE's own Database (src/control/database.cpp compiled where it lies over the shims in
oracle/shim -> oracle/_ref/libsmath_ref.so, entry points oracle/ref_db_wrap.cpp): volumes, exclusion factors and scattering
factors of every function type (database.cpp:391-528, with the reference's powf / float-sqrt roundings), the effective
scattering length of ScatterFactors::update (scatter_factors.cpp:56-78) and the PDB atom-name resolution (:309-340), for the
test database of tests/test_control_plane.py.  Run in the build container (needs /root/reference); writes
tests/golden/ref_database.npz, which travels with the repo.

    python tests/golden/make_ref_database_golden.py
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as o  # noqa: E402
import test_control_plane as t  # noqa: E402

subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
t.register_reference_database(o)
els = t.REF_DB_ELEMENTS
q = np.array([0.0, 0.05, 0.4, 0.7, 1.1, 1.6, 2.3, 5.0])
kappa = np.array([1.0, 1.5, 0.25])
bg = 0.0334
vol = np.array([o.ref_db_volume(i) for i in range(len(els))])
sf = np.array([[o.ref_db_sfactor(i, x) for x in q] for i in range(len(els))])
ex = np.array([[[o.ref_db_exclusion(i, k * vol[i], x) for x in q] for k in kappa] for i in range(len(els))])
eff = np.array([[[o.ref_db_effective(i, x, k, bg) for x in q] for k in kappa] for i in range(len(els))])
names = sorted(t.NAME2EL)
resolved = np.array([o.ref_db_name_get(n) for n in names])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_database.npz"), elements=np.array(els), q=q, kappa=kappa,
                    background=bg, volume=vol, sfactor=sf, exclusion=ex, effective=eff, names=np.array(names), resolved=resolved)
print("wrote tests/golden/ref_database.npz")
