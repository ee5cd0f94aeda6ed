/* boost/algorithm/string.hpp — SHIM: trim() (in place, whitespace on both sides), the one function the reference's selection
 * readers use */
#ifndef ORACLE_SHIM_BOOST_ALGORITHM_STRING_HPP
#define ORACLE_SHIM_BOOST_ALGORITHM_STRING_HPP
#include <cctype>
#include <string>
namespace boost {
inline void trim(std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    s = s.substr(a, b - a);
}
}  // namespace boost
#endif
