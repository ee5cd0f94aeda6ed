// ref_params_wrap.cpp — calls the three generators of the REFERENCE's own src/control/parameters.cpp (:930-1189: orientation
// vectors, multipole moments, q-vector scans), compiled where it lies into oracle/_ref/libparams_ref.so over the shims in
// oracle/shim_params (program_options, filesystem, an XMLInterface that is never called: libxml2 is absent) and oracle/shim
// (Boost.Random: a Boost-1.4x restatement, so the sphere / cylinder "boost_uniform_on_sphere" algorithms pin the USE of the
// stream, not the stream).  A library of its own: the real Params class would clash with the plain-struct Params the scatter
// devices in libsmath_ref.so are built against.  Test infrastructure.
#include <string>
#include <vector>

#include "control.hpp"

extern "C" {
// type "sphere" | "cylinder" | "file"; returns the number of vectors; out may be NULL to query the count
size_t ref_orientation_vectors(const char *type, const char *algorithm, size_t resolution, unsigned long seed, const char *filepath,
                               double *out, size_t cap) {
    ScatteringAverageOrientationVectorsParameters v;
    v.type = type;
    v.algorithm = algorithm;
    v.resolution = resolution;
    v.seed = seed;
    v.filepath = filepath;
    v.create();
    for (size_t i = 0; i < v.size() && i < cap && out; i++) {
        out[3 * i] = v[i].x;
        out[3 * i + 1] = v[i].y;
        out[3 * i + 2] = v[i].z;
    }
    return v.size();
}
// multipole_type "sphere" | "cylinder"; type "resolution" | "file"
size_t ref_multipole_moments(const char *multipole_type, const char *type, long resolution, const char *filepath, long *out, size_t cap) {
    Params::Inst()->scattering.average.orientation.multipole.type = multipole_type;
    ScatteringAverageOrientationMultipoleMomentsParameters m;
    m.type = type;
    m.resolution = resolution;
    m.filepath = filepath;
    m.create();
    for (size_t i = 0; i < m.size() && i < cap && out; i++) {
        out[2 * i] = m[i].first;
        out[2 * i + 1] = m[i].second;
    }
    return m.size();
}
size_t ref_scan_vectors(size_t nscans, const double *from, const double *to, const size_t *points, const double *exponent,
                        const double *base, double *out, size_t cap) {
    ScatteringVectorsParameters q;
    for (size_t i = 0; i < nscans; i++) {
        ScatteringVectorsScanParameters s;
        s.from = from[i];
        s.to = to[i];
        s.points = points[i];
        s.exponent = exponent[i];
        s.basevector = CartesianCoor3D(base[3 * i], base[3 * i + 1], base[3 * i + 2]);
        q.scans.push_back(s);
    }
    q.create_from_scans();
    for (size_t i = 0; i < q.size() && i < cap && out; i++) {
        out[3 * i] = q[i].x;
        out[3 * i + 1] = q[i].y;
        out[3 * i + 2] = q[i].z;
    }
    return q.size();
}
}
