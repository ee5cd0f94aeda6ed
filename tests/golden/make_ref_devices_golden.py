"""Golden vectors produced by the REFERENCE's own scatter devices: AllVectorsScatterDevice and SelfVectorsScatterDevice (with
the reference's stagers, DSP and partition code) compiled where they lie for one MPI rank over the shims in oracle/shim
(make -C oracle ref -> oracle/_ref/libsmath_ref.so, entry point oracle/ref_devices_wrap.cpp).  Run in the build container
(needs /root/reference); writes tests/golden/ref_devices.npz, which travels with the repo and pins the oracle AND the CUDA
path on machines without the reference.

    python tests/golden/make_ref_devices_golden.py
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402
from sassena_b200 import synth  # noqa: E402

subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
out = {}
cases = []
NF, NA = 37, 53
xyz = synth.trajectory(NF, NA, 25.0, 0.3, 101, offset=-12.5)
b = synth.factors(NA)
out["xyz"], out["b"] = xyz, b
qv = np.array([[0.7, 0.0, 0.0], [0.0, 1.3, 0.2], [-0.4, 0.4, 2.0]])
out["qv"] = qv
u = synth.unit_vectors(11, 2)
out["u"] = u
phi = np.linspace(0, 2 * np.pi, 9, endpoint=False)
cyl = np.stack([np.cos(phi), np.sin(phi), np.zeros_like(phi)], axis=1)
out["cyl"] = cyl
axis = np.array([1.0, -1.0, 2.0])
out["axis"] = axis
k = 0
for kind in ("all", "self"):
    for vt, ori in (("file", u), ("sphere", u), ("cylinder", cyl), ("none", None)):
        for dsp, method in (("autocorrelate", "fftw"), ("autocorrelate", "direct"), ("square", "fftw"), ("plain", "fftw")):
            if vt in ("sphere", "none") and (dsp, method) != ("autocorrelate", "fftw"):
                continue
            q, fqt, fq, fq2 = o.ref_scatter_run(kind, xyz, b, qv, orient=ori, vectors_type="file" if vt == "none" else vt, axis=axis,
                                                dsp=dsp, method=method, threads=2)
            assert np.array_equal(q, qv)
            out[f"case{k}_fqt"], out[f"case{k}_fq"], out[f"case{k}_fq2"] = fqt, fq, fq2
            cases.append((kind, vt, dsp, method))
            k += 1
out["cases"] = np.array(cases)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_devices.npz"), **out)
print("wrote tests/golden/ref_devices.npz:", k, "cases")
