/* sample.hpp — SHIM: the reference's Sample (atoms + selections + framesets + alignments/motions, Boost + libxml2) reduced to
 * what the stagers and scatter devices read: the size of the target selection and coordinate_sets.load(frame), served from
 * arrays the test harness provides.  The representation change uses the reference's own SphericalCoor3D / CylinderCoor3D. */
#ifndef ORACLE_SHIM_SAMPLE_HPP
#define ORACLE_SHIM_SAMPLE_HPP
#include <cstddef>
#include <map>
#include <string>
#include <vector>
#include "math/coor3d.hpp"
enum CoordinateRepresentation { CARTESIAN = 10, SPHERICAL = 20, CYLINDRICAL = 30 };
class IAtomselection {
   public:
    virtual ~IAtomselection() {}
    virtual size_t operator[](size_t index) = 0;
    virtual size_t size() = 0;
};
class ShimRangeSelection : public IAtomselection {
    size_t n_;
   public:
    explicit ShimRangeSelection(size_t n) : n_(n) {}
    size_t operator[](size_t i) { return i; }
    size_t size() { return n_; }
};
class CoordinateSet {
   public:
    std::vector<coor2_t> c1, c2, c3;
    size_t size() { return c1.size(); }
};
class CoordinateSets {
    const float *frames_ = nullptr;  // [NF][NA][3] cartesian, the target selection
    size_t NF_ = 0, NA_ = 0;
    CoordinateRepresentation repr_ = CARTESIAN;
    CartesianCoor3D axis_;
   public:
    void shim_set(const float *frames, size_t NF, size_t NA, CartesianCoor3D axis) { frames_ = frames; NF_ = NF; NA_ = NA; axis_ = axis; }
    size_t size() { return NF_; }
    void set_selection(IAtomselection *) {}
    void set_representation(CoordinateRepresentation r) { repr_ = r; }
    CoordinateRepresentation get_representation() { return repr_; }
    CoordinateSet *load(size_t f) {  // CoordinateSets::load (coordinate_sets.cpp:343-357) without alignments / motions
        CoordinateSet *cs = new CoordinateSet;
        CartesianVectorBase base(axis_);
        for (size_t n = 0; n < NA_; n++) {
            CartesianCoor3D c(frames_[(f * NA_ + n) * 3], frames_[(f * NA_ + n) * 3 + 1], frames_[(f * NA_ + n) * 3 + 2]);
            if (repr_ == SPHERICAL) {
                SphericalCoor3D s(c);
                cs->c1.push_back(s.r); cs->c2.push_back(s.phi); cs->c3.push_back(s.theta);
            } else if (repr_ == CYLINDRICAL) {
                CylinderCoor3D y(base.project(c));
                cs->c1.push_back(y.r); cs->c2.push_back(y.phi); cs->c3.push_back(y.z);
            } else {
                cs->c1.push_back(c.x); cs->c2.push_back(c.y); cs->c3.push_back(c.z);
            }
        }
        return cs;
    }
};
class Atoms {
   public:
    std::map<std::string, IAtomselection *> selections;
};
class Sample {
   public:
    Atoms atoms;
    CoordinateSets coordinate_sets;
};
#endif
