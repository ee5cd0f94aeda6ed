/* boost/filesystem.hpp — SHIM: exists() only (database.cpp checks that the database file is there) */
#ifndef ORACLE_SHIM_BOOST_FILESYSTEM_HPP
#define ORACLE_SHIM_BOOST_FILESYSTEM_HPP
#include <fstream>
#include <string>
namespace boost {
namespace filesystem {
inline bool exists(const std::string &p) { return std::ifstream(p.c_str()).good(); }
}  // namespace filesystem
}  // namespace boost
#endif
