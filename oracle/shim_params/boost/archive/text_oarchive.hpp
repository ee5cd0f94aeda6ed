/* boost/archive/text_oarchive.hpp — SHIM: lets frameset_index.hpp compile.  Boost.Serialization is absent, so the reference's .tnx
 * frame-index files (Boost text archives) cannot be produced or read in oracle/_ref: load()/save() throw.  The readers are
 * exercised through generate_index() instead. */
#ifndef ORACLE_SHIM_BOOST_ARCHIVE_TEXT_OARCHIVE_HPP
#define ORACLE_SHIM_BOOST_ARCHIVE_TEXT_OARCHIVE_HPP
#include <istream>
#include <ostream>
#include <stdexcept>
namespace boost {
namespace archive {
class text_oarchive {
   public:
    template <class S> explicit text_oarchive(S &) {}
    template <class T> text_oarchive &operator<<(const T &) { throw std::runtime_error("oracle shim: Boost text archives are not available"); }
    template <class T> text_oarchive &operator>>(T &) { throw std::runtime_error("oracle shim: Boost text archives are not available"); }
    template <class T> text_oarchive &operator&(T &) { throw std::runtime_error("oracle shim: Boost text archives are not available"); }
};
}  // namespace archive
}  // namespace boost
#endif
