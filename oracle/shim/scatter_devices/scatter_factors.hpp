/* scatter_devices/scatter_factors.hpp — SHIM: b_j(|q|) comes from the Database singleton in the reference (db.xml, Boost +
 * libxml2); here update(q) asks a callback of the test harness for the factors of the target selection. */
#ifndef ORACLE_SHIM_SCATTER_FACTORS_HPP
#define ORACLE_SHIM_SCATTER_FACTORS_HPP
#include <cstddef>
#include <vector>
#include "math/coor3d.hpp"
#include "sample.hpp"
typedef void (*shim_factors_cb)(void *user, double ql, double *b, size_t n);
struct ShimFactorSource {
    shim_factors_cb cb = nullptr;
    void *user = nullptr;
    static ShimFactorSource &Inst() {
        static ShimFactorSource s;
        return s;
    }
};
class ScatterFactors {
    std::vector<double> f_;
    size_t n_ = 0;
   public:
    void set_sample(Sample &) {}
    void set_selection(IAtomselection *s) { n_ = s->size(); f_.assign(n_, 0.0); }
    void set_background(bool) {}
    void update(CartesianCoor3D q) { if (ShimFactorSource::Inst().cb) ShimFactorSource::Inst().cb(ShimFactorSource::Inst().user, q.length(), f_.data(), n_); }
    double get(size_t i) { return f_[i]; }
    std::vector<double> &get_all() { return f_; }
    double compute_background(CartesianCoor3D) { return 0.0; }
};
#endif
