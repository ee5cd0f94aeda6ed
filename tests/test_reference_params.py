"""The q-vector / orientation / moment generators against the REFERENCE's own parameters.cpp (:930-1189), compiled where it lies
into oracle/_ref/libparams_ref.so (oracle/Makefile; libxml2, program_options, filesystem stood in for by oracle/shim_params).
tests/golden/ref_params.npz holds its output (tests/golden/make_ref_params_golden.py); the oracle and the product's host layer
(csrc/host/sassena_host.cpp through sassena_b200.host) reproduce every array bit for bit -- these generators define which
q-vectors the hot path sees.  The uniform_on_sphere cases run over a Boost-1.4x restatement of Boost.Random
(absent here): they pin the use of the stream, not the stream."""
import os

import numpy as np
import pytest

from sassena_b200 import host

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_params.npz")
SEEDS = (0, 7, 4357)
NV = 60
RASTER = (1, 2, 5)
MOMENT_RES = (0, 1, 4, 20)
FILE_VECTORS = np.array([[2.0, 0.0, 0.0], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [-0.3, 0.4, 1.2]])
FILE_MOMENTS = np.array([[0, 0], [2, -1], [3, 3], [5, 0]])
SCANS = [
    [{"base": (1, 0, 0), "from": 0.2, "to": 2.0, "points": 10}],
    [{"base": (0.6, 0.0, 0.8), "from": 0.1, "to": 5.0, "points": 50, "exponent": 1.0}],  # BASELINE config 3: 50 |q|
    [{"base": (1, 0, 0), "from": 0, "to": 1, "points": 3}, {"base": (0, 1, 0), "from": 0, "to": 2, "points": 2},
     {"base": (0, 0, 1), "from": 1, "to": 3, "points": 1, "exponent": 2.0}],
    [{"base": (1, 1, 0), "from": 0.1, "to": 3.0, "points": 17, "exponent": 2.5},
     {"base": (0, 0, 1), "from": -1, "to": 1, "points": 5, "exponent": 0.5}],
    [{"base": (0, 0, 1), "from": 0.5, "to": 1.5, "points": 0}, {"base": (0, 1, 0), "from": 0.5, "to": 1.5, "points": 4}],
]


def _params_vectors(vtype, algorithm="boost_uniform_on_sphere", resolution=NV, seed=0):
    p = host.Params().set("scattering.average.orientation.type", "vectors")
    p.set("scattering.average.orientation.vectors.type", vtype).set("scattering.average.orientation.vectors.algorithm", algorithm)
    p.set("scattering.average.orientation.vectors.resolution", resolution).set("scattering.average.orientation.vectors.seed", seed)
    return p


def test_oracle_generators_equal_reference(oracle):
    g = np.load(GOLD)
    for seed in SEEDS:
        assert np.array_equal(oracle.uniform_on_sphere(seed, 3, NV), g[f"sphere_{seed}"])
        assert np.array_equal(oracle.uniform_on_sphere(seed, 2, NV), g[f"cylinder_{seed}"])
    for r in RASTER:
        assert np.array_equal(oracle.cylinder_raster_linear(r), g[f"raster_{r}"])
    for L in MOMENT_RES:
        assert np.array_equal(oracle.moments_sphere(L), g[f"moments_sphere_{L}"])
        assert np.array_equal(oracle.moments_cylinder(L), g[f"moments_cylinder_{L}"])
    for i, s in enumerate(SCANS):
        assert np.array_equal(np.reshape(oracle.qvectors_from_scans(s), (-1, 3)), g[f"scan_{i}"]), i  # (scan 4 is empty)


def test_product_generators_equal_reference():
    g = np.load(GOLD)
    for seed in SEEDS:
        assert np.array_equal(_params_vectors("sphere", seed=seed).create().vectors, g[f"sphere_{seed}"])
        assert np.array_equal(_params_vectors("cylinder", seed=seed).create().vectors, g[f"cylinder_{seed}"])
    for r in RASTER:
        assert np.array_equal(_params_vectors("cylinder", "raster_linear", r).create().vectors, g[f"raster_{r}"])
    f = _params_vectors("file").set_vectors(FILE_VECTORS).create().vectors
    assert np.array_equal(f, g["vectors_file"])
    for mtype in ("sphere", "cylinder"):
        for L in MOMENT_RES:
            p = host.Params().set("scattering.average.orientation.type", "multipole")
            p.set("scattering.average.orientation.multipole.type", mtype)
            p.set("scattering.average.orientation.multipole.moments.type", "resolution")
            p.set("scattering.average.orientation.multipole.moments.resolution", L).create()
            assert np.array_equal(p.moments, g[f"moments_{mtype}_{L}"]), (mtype, L)
    p = host.Params().set("scattering.average.orientation.type", "multipole")
    p.set("scattering.average.orientation.multipole.moments.type", "file").set_moments(FILE_MOMENTS).create()
    assert np.array_equal(p.moments, g["moments_file"])
    for i, s in enumerate(SCANS):
        assert np.array_equal(np.reshape(host.create_from_scans(s), (-1, 3)), g[f"scan_{i}"]), i


def test_reference_generators_live(oracle, tmp_path):
    if not oracle.have_ref_params():
        pytest.skip("oracle/_ref/libparams_ref.so not built (no /root/reference on this machine)")
    s = [{"base": (0.0, 1.0, 0.0), "from": 0.3, "to": 4.1, "points": 23, "exponent": 1.7}]
    assert np.array_equal(oracle.ref_scan_vectors(s), host.create_from_scans(s))
    assert np.array_equal(oracle.ref_scan_vectors(s), oracle.qvectors_from_scans(s))
    assert np.array_equal(oracle.ref_orientation_vectors("sphere", resolution=33, seed=99), oracle.uniform_on_sphere(99, 3, 33))
    assert np.array_equal(oracle.ref_multipole_moments("sphere", 9), oracle.moments_sphere(9))
    path = str(tmp_path / "qqq.txt")
    np.savetxt(path, FILE_VECTORS)
    assert np.array_equal(oracle.ref_orientation_vectors("file", filepath=path),
                          _params_vectors("file").set_vectors(FILE_VECTORS).create().vectors)
