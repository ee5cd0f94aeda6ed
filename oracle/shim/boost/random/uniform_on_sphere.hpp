/* boost/random/uniform_on_sphere.hpp — SHIM restating Boost 1.4x: dim normal variates, scaled by 1/sqrt(sum of squares) */
#ifndef ORACLE_SHIM_BOOST_UOS_HPP
#define ORACLE_SHIM_BOOST_UOS_HPP
#include <cmath>
#include <vector>
#include <boost/random/normal_distribution.hpp>
namespace boost {
template <class RealType = double, class Cont = std::vector<RealType> >
class uniform_on_sphere {
    normal_distribution<RealType> normal_;
    Cont container_;
    int dim_;
   public:
    typedef Cont result_type;
    explicit uniform_on_sphere(int dim = 2) : container_(dim), dim_(dim) {}
    template <class Engine>
    const result_type &operator()(Engine &eng) {
        RealType sqsum = 0;
        for (int i = 0; i < dim_; i++) {
            RealType val = normal_(eng);
            container_[i] = val;
            sqsum += val * val;
        }
        const RealType inv = RealType(1) / std::sqrt(sqsum);
        for (int i = 0; i < dim_; i++) container_[i] *= inv;
        return container_;
    }
};
}  // namespace boost
#endif
