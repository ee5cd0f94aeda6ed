/* boost/random/variate_generator.hpp — SHIM restating Boost 1.4x for an integer engine and a real distribution: the engine is
 * copied and wrapped as uniform_01, (x - min) / (max - min + 1) = x / 2^32 for mt19937 */
#ifndef ORACLE_SHIM_BOOST_VARGEN_HPP
#define ORACLE_SHIM_BOOST_VARGEN_HPP
namespace boost {
template <class Engine, class Distribution>
class variate_generator {
    struct U01 {
        Engine e;
        explicit U01(const Engine &x) : e(x) {}
        double operator()() { return (double)(e() - (Engine::min)()) / ((double)((Engine::max)() - (Engine::min)()) + 1.0); }
    } u_;
    Distribution d_;
   public:
    typedef typename Distribution::result_type result_type;
    variate_generator(Engine e, Distribution d) : u_(e), d_(d) {}
    result_type operator()() { return d_(u_); }
};
}  // namespace boost
#endif
