/* boost/random/normal_distribution.hpp — SHIM restating Boost 1.4x: Box-Muller, one cached value; the engine it is handed
 * yields uniform_01 variates (see variate_generator.hpp).  BEST EFFORT: later Boost versions draw differently. */
#ifndef ORACLE_SHIM_BOOST_NORMAL_HPP
#define ORACLE_SHIM_BOOST_NORMAL_HPP
#include <cmath>
namespace boost {
template <class RealType = double>
class normal_distribution {
    RealType mean_, sigma_, r1_, r2_, cached_rho_;
    bool valid_;
   public:
    typedef RealType result_type;
    explicit normal_distribution(RealType mean = RealType(0), RealType sigma = RealType(1))
        : mean_(mean), sigma_(sigma), r1_(0), r2_(0), cached_rho_(0), valid_(false) {}
    template <class Engine>
    result_type operator()(Engine &eng) {
        if (!valid_) {
            r1_ = eng();
            r2_ = eng();
            cached_rho_ = std::sqrt(-RealType(2) * std::log(RealType(1) - r2_));
            valid_ = true;
        } else {
            valid_ = false;
        }
        const RealType pi = RealType(3.14159265358979323846);
        return cached_rho_ * (valid_ ? std::cos(RealType(2) * pi * r1_) : std::sin(RealType(2) * pi * r1_)) * sigma_ + mean_;
    }
};
}  // namespace boost
#endif
