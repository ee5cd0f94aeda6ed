// ref_smath_wrap.cpp — C entry points into the REFERENCE's own smath.cpp (compiled where it lies, see oracle/Makefile
// target `ref`) plus the FFTW3 shim's implementation.  Test infrastructure: tests/test_oracle.py checks the oracle's
// restatement of smath.cpp:51-76,141-156,168-177 against these.
#include <cstdlib>
#include <cstring>
#include <complex>
#include <vector>

#include <fftw3.h>

#include "math/smath.hpp"

extern "C" void orc_fft(double *data, size_t n, int sign);  // oracle/sassena_oracle.c (linked in)

extern "C" {
void *fftw_malloc(size_t n) { return std::malloc(n); }
void fftw_free(void *p) { std::free(p); }
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *, fftw_complex *, int sign, unsigned) {
    fftw_plan p = static_cast<fftw_plan>(std::malloc(sizeof(*p)));
    p->n = n;
    p->sign = sign;
    return p;
}
void fftw_destroy_plan(fftw_plan p) { std::free(p); }
void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out) {
    if (out != in) std::memcpy(out, in, sizeof(fftw_complex) * p->n);
    orc_fft(&out[0][0], (size_t)p->n, p->sign);
}

// smath::auto_correlate_direct(fftw_complex*, N)                              smath.cpp:51-76
void ref_auto_correlate_direct(double *data, size_t NF) { smath::auto_correlate_direct(reinterpret_cast<fftw_complex *>(data), NF); }
// smath::auto_correlate_fftw(fftw_complex*, p1, p2, NF) on a 2NF buffer whose upper half the caller zeroed, with the plans
// the devices make (all_vectors_scatter_device.cpp:52-53)                      smath.cpp:141-156
void ref_auto_correlate_fftw(double *data2nf, size_t NF) {
    fftw_complex *d = reinterpret_cast<fftw_complex *>(data2nf);
    fftw_plan p1 = fftw_plan_dft_1d((int)(2 * NF), d, d, FFTW_FORWARD, FFTW_ESTIMATE);
    fftw_plan p2 = fftw_plan_dft_1d((int)(2 * NF), d, d, FFTW_BACKWARD, FFTW_ESTIMATE);
    smath::auto_correlate_fftw(d, p1, p2, NF);
    fftw_destroy_plan(p1);
    fftw_destroy_plan(p2);
}
// the std::vector overloads (unused by the devices; their direct form conjugates the other way round)  smath.cpp:24-48,100-139
void ref_auto_correlate_direct_vec(double *data, size_t NF) {
    std::vector<std::complex<double> > v(NF);
    std::memcpy(v.data(), data, sizeof(double) * 2 * NF);
    smath::auto_correlate_direct(v);
    std::memcpy(data, v.data(), sizeof(double) * 2 * NF);
}
// smath::square_elements(fftw_complex*, N)                                     smath.cpp:168-177
void ref_square_elements(double *data, size_t NF) { smath::square_elements(reinterpret_cast<fftw_complex *>(data), NF); }
}

// ---- the reference's own src/math/coor3d.cpp and src/decomposition/assignment.cpp ------------------------------------------
#include "decomposition/assignment.hpp"
#include "math/coor3d.hpp"

extern "C" {
// SphericalCoor3D(CartesianCoor3D) (coor3d.cpp:168-215): out = (r, phi, theta)
void ref_cart_to_spherical(const double xyz[3], double out[3]) {
    SphericalCoor3D s(CartesianCoor3D(xyz[0], xyz[1], xyz[2]));
    out[0] = s.r;
    out[1] = s.phi;
    out[2] = s.theta;
}
// CylinderCoor3D(CartesianVectorBase(axis).project(c)) (coor3d.cpp:113-138,278-304): out = (r, phi, z)
void ref_cart_to_cylindrical(const double xyz[3], const double axis[3], double out[3]) {
    CartesianVectorBase base(CartesianCoor3D(axis[0], axis[1], axis[2]));
    CylinderCoor3D c(base.project(CartesianCoor3D(xyz[0], xyz[1], xyz[2])));
    out[0] = c.r;
    out[1] = c.phi;
    out[2] = c.z;
}
// the base itself: e_r, e_phi, e_z as rows
void ref_vector_base(const double axis[3], double out[9]) {
    CartesianVectorBase base(CartesianCoor3D(axis[0], axis[1], axis[2]));
    for (int i = 0; i < 3; i++) {
        out[3 * i] = base[i].x;
        out[3 * i + 1] = base[i].y;
        out[3 * i + 2] = base[i].z;
    }
}
// DivAssignment / ModAssignment (assignment.cpp:27-132): offset, size, max and the first `cap` indices
void ref_assignment(int mod, size_t NN, size_t rank, size_t NAF, size_t *offset, size_t *size, size_t *max, size_t *idx, size_t cap) {
    if (mod) {
        ModAssignment a(NN, rank, NAF);
        *offset = a.offset();
        *size = a.size();
        *max = a.max();
        for (size_t i = 0; i < a.size() && i < cap; i++) idx[i] = a[i];
    } else {
        DivAssignment a(NN, rank, NAF);
        *offset = a.offset();
        *size = a.size();
        *max = a.max();
        for (size_t i = 0; i < a.size() && i < cap; i++) idx[i] = a[i];
    }
}
}

// ---- the reference's own src/decomposition/decomposition_plan.cpp -------------------------------------------------------------
#include "control.hpp"
#include "decomposition/decomposition_plan.hpp"

extern "C" {
// DecompositionParameters(NN, NQ, NAF, NNpP, elbytesize).penalty()                 decomposition_plan.cpp:28-66
size_t ref_decomposition_penalty(size_t NN, size_t NQ, size_t NAF, size_t NNpP) { return DecompositionParameters(NN, NQ, NAF, NNpP, 1).penalty(); }
// DecompositionPlan(nn, nq, naf, elbytesize, nmaxbytesize) with the three Params values it reads.  ONLY for inputs that have a
// valid plan: the reference reports failure with a bare `throw;`, which terminates the process.    decomposition_plan.cpp:69-157
void ref_decomposition_plan(size_t nn, size_t nq, size_t naf, size_t elbytes, size_t maxbytes, int automatic, size_t manual_size,
                            double utilization, size_t *partitions, size_t *partitionsize, size_t *penalty, size_t *colors) {
    Params::Inst()->limits.decomposition.partitions.automatic = automatic != 0;
    Params::Inst()->limits.decomposition.partitions.size = manual_size;
    Params::Inst()->limits.decomposition.utilization = utilization;
    DecompositionPlan dp(nn, nq, naf, elbytes, maxbytes);
    *partitions = dp.partitions();
    *partitionsize = dp.partitionsize();
    *penalty = DecompositionParameters(nn, nq, naf, dp.partitionsize(), elbytes).penalty();  // (DecompositionPlan::penalty() is declared but not defined in the reference)
    std::vector<size_t> c = dp.colors();
    for (size_t i = 0; i < c.size(); i++) colors[i] = c[i];
}
}

// ---- the reference's own src/stager/coordinate_writer.cpp (DCD layout of stager.dump) ------------------------------------------
#include "stager/coordinate_writer.hpp"

extern "C" {
// DCDCoordinateWriter(file, blocks, entries): init(); prepare(); write(data, 0, blocks)      coordinate_writer.cpp:30-144
// data: float [blocks][entries][3] (coor_t = float)
void ref_dcd_write(const char *path, float *data, size_t blocks, size_t entries) {
    DCDCoordinateWriter w(path, blocks, entries);
    w.init();
    w.prepare();
    w.write(data, 0, blocks);
}
// the same file written in two pieces the way a 2-rank partition does (data_stager.cpp:148-160): blocks [0, split) and [split, blocks)
void ref_dcd_write_split(const char *path, float *data, size_t blocks, size_t entries, size_t split) {
    DCDCoordinateWriter w0(path, blocks, entries), w1(path, blocks, entries);
    w0.init();
    w0.prepare();
    w1.prepare();
    w1.write(data + split * entries * 3, split, blocks - split);
    w0.write(data, 0, split);
}
}

// ---- the reference's own src/sample/motion_walker.cpp (over the uBLAS / Boost.Random shims) -----------------------------------
#include "sample/motion_walker.hpp"

extern "C" {
// transform(t) of a walker for t = 0 .. n-1: out[t] = the 4x4 matrix, row major.  type: 0 linear, 1 fixed, 2 oscillation,
// 3 randomwalk, 4 brownian, 5 localbrownian, 6 rotationalbrownian (coordinate_sets.cpp:120-148 picks them by name)
int ref_motion_transforms(int type, double displace, double frequency, double radius, unsigned long seed, long sampling,
                          const double dir[3], size_t n, double *out) {
    CartesianCoor3D d(dir[0], dir[1], dir[2]);
    MotionWalker *w = NULL;
    switch (type) {
        case 0: w = new LinearMotionWalker(displace, sampling, d); break;
        case 1: w = new FixedMotionWalker(displace, d); break;
        case 2: w = new OscillationMotionWalker(displace, frequency, sampling, d); break;
        case 3: w = new RandomMotionWalker(displace, seed, sampling, d); break;
        case 4: w = new BrownianMotionWalker(displace, seed, sampling, d); break;
        case 5: w = new LocalBrownianMotionWalker(radius, displace, seed, sampling, d); break;
        case 6: w = new RotationalBrownianMotionWalker(displace, seed, sampling); break;
        default: return 1;
    }
    for (size_t t = 0; t < n; t++) {
        boost::numeric::ublas::matrix<double> T = w->transform(t);
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) out[16 * t + 4 * i + j] = T(i, j);
    }
    delete w;
    return 0;
}
}
