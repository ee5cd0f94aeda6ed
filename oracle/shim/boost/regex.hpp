/* boost/regex.hpp — SHIM over <regex> (ECMAScript grammar; the database's atom-name patterns and the ".*\.xml$" test use the
 * subset on which Boost's Perl grammar and ECMAScript agree) */
#ifndef ORACLE_SHIM_BOOST_REGEX_HPP
#define ORACLE_SHIM_BOOST_REGEX_HPP
#include <regex>
#include <string>
namespace boost {
typedef std::regex regex;
inline bool regex_match(const std::string &s, const regex &e) { return std::regex_match(s, e); }
}  // namespace boost
#endif
