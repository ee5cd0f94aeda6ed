"""Golden vectors produced by the REFERENCE's own multipole devices: MPSphereScatterDevice and MPCylinderScatterDevice
(src/scatter_devices/multipole_scatter_device.cpp, with the reference's frame stager, coor3d conversions, DSP and store) compiled
where they lie for one MPI rank over the shims in oracle/shim (make -C oracle ref -> oracle/_ref/libsmath_ref.so, entry point
ref_multipole_run in oracle/ref_devices_wrap.cpp).  Boost.Math is not in this image: sph_bessel / spherical_harmonic /
cyl_bessel_j are served by the oracle's restatements (oracle/shim/boost/math/special_functions.hpp), so these vectors pin the
devices' USE of the functions -- moment order, prefactors, conjugation, summation order, normalisation, dsp, store, final
scaling -- not the functions' last bits (tests/test_oracle.py checks those against scipy).  Run in the build container (needs
/root/reference); writes tests/golden/ref_multipole_devices.npz, which travels with the repo.

    python tests/golden/make_ref_multipole_golden.py
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
NF, NA = 23, 61
xyz = synth.trajectory(NF, NA, 22.0, 0.4, 211, offset=-11.0)
b = synth.factors(NA)
out["xyz"], out["b"] = xyz, b
qv = np.array([[0.35, 0.0, 0.0], [0.0, -0.9, 0.4], [1.2, 0.7, -1.6]])
out["qv"] = qv
axis = np.array([0.3, -0.2, 1.0])
out["axis"] = axis
out["mom_sphere"] = o.moments_sphere(9)
out["mom_cylinder"] = o.moments_cylinder(6)
k = 0
for kind in ("sphere", "cylinder"):
    for dsp, method in (("autocorrelate", "fftw"), ("autocorrelate", "direct"), ("square", "fftw"), ("plain", "fftw")):
        # worker threads: a divisor of the number of moments (100 / 25) -- the reference's padding of the last block of moments
        # makes its workers read multipole_index_[NM] out of bounds (multipole_scatter_device.cpp:199,418)
        threads = 2 if kind == "sphere" else 5
        assert len(out["mom_" + kind]) % threads == 0
        q, fqt, fq, fq2 = o.ref_multipole_run(kind, xyz, b, qv, out["mom_" + kind], axis=axis, dsp=dsp, method=method, threads=threads)
        assert np.array_equal(q, qv)
        out[f"case{k}_fqt"], out[f"case{k}_fq"], out[f"case{k}_fq2"] = fqt, fq, fq2
        cases.append((kind, dsp, method))
        k += 1
out["cases"] = np.array(cases)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_multipole_devices.npz"), **out)
print("wrote tests/golden/ref_multipole_devices.npz:", k, "cases")
