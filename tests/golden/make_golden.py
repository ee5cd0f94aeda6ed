"""Generates tests/golden/*.npz: small seeded inputs and the oracle's outputs for them.

The reference itself cannot run in this image (needs Boost/FFTW3/MPI/HDF5), so these vectors come from the CPU
oracle (oracle/sassena_oracle.c) after it has been cross-checked against numpy/scipy and the analytic known
answers in tests/test_oracle.py.  They pin the oracle AND the CUDA path against silent drift.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as o  # noqa: E402
from sassena_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # coherent: 40 atoms, 24 frames, 3 |q| x 9 vectors, all dsp modes
    NA, NF = 40, 24
    xyz = synth.trajectory(NF, NA, 25.0, 0.3, 101)
    b = synth.factors(NA)
    u = synth.unit_vectors(9, 102)
    qls = synth.qlengths(0.3, 1.9, 3)
    out = {"xyz": xyz, "b": b, "u": u, "qls": qls}
    for dsp, method in (("autocorrelate", "fftw"), ("autocorrelate", "direct"), ("square", "fftw"), ("plain", "fftw")):
        for i, ql in enumerate(qls):
            fqt, fq, fq2 = o.compute_all_vectors(xyz, b, ql * u, dsp=dsp, method=method)
            out[f"all_{dsp}_{method}_{i}_fqt"] = fqt
            out[f"all_{dsp}_{method}_{i}_fq"] = np.array([fq, fq2])
    np.savez_compressed(os.path.join(HERE, "coherent_small.npz"), **out)

    # self: 12 atoms, 20 frames, 2 |q| x 5 vectors
    NA, NF = 12, 20
    xa = synth.trajectory(NF, NA, 25.0, 0.3, 103, layout=1)
    b = synth.factors(NA)
    u = synth.unit_vectors(5, 104)
    out = {"xyz_by_atom": xa, "b": b, "u": u, "qls": np.array([0.5, 1.5])}
    for i, ql in enumerate((0.5, 1.5)):
        fqt, fq, fq2 = o.compute_self_vectors(xa, b, ql * u)
        out[f"self_{i}_fqt"] = fqt
        out[f"self_{i}_fq"] = np.array([fq, fq2])
    np.savez_compressed(os.path.join(HERE, "self_small.npz"), **out)

    # multipole sphere: 25 atoms, 6 frames, L=5, two |q|
    NA, NF = 25, 6
    xyz = synth.trajectory(NF, NA, 30.0, 0.3, 105, offset=-15.0)
    b = synth.factors(NA)
    mom = o.moments_sphere(5)
    out = {"xyz": xyz, "b": b, "moments": mom, "qls": np.array([0.1, 0.9])}
    for i, ql in enumerate((0.1, 0.9)):
        fqt, fq, fq2 = o.compute_mpsphere(o.cart_to_spherical(xyz), b, ql, mom)
        out[f"mp_{i}_fqt"] = fqt
        out[f"mp_{i}_fq"] = np.array([fq, fq2])
    np.savez_compressed(os.path.join(HERE, "mpsphere_small.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
