/* boost/lexical_cast.hpp — SHIM: src/decomposition/assignment.cpp formats numbers into its error messages with it. */
#ifndef ORACLE_SHIM_BOOST_LEXICAL_CAST_HPP
#define ORACLE_SHIM_BOOST_LEXICAL_CAST_HPP
#include <sstream>
#include <string>
namespace boost {
template <class Target, class Source>
Target lexical_cast(const Source &s) {
    std::stringstream ss;
    ss << s;
    Target t;
    ss >> t;
    return t;
}
}  // namespace boost
#endif
