/* boost/filesystem.hpp — SHIM over std::filesystem for the build of the reference's parameters.cpp (Params::get_filepath and
 * the default .tnx index name) */
#ifndef ORACLE_SHIM_PARAMS_BOOST_FILESYSTEM_HPP
#define ORACLE_SHIM_PARAMS_BOOST_FILESYSTEM_HPP
#include <filesystem>
#include <string>
namespace boost {
namespace filesystem {
class path {
    std::filesystem::path p_;
   public:
    path() {}
    path(const std::string &s) : p_(s) {}
    path(const char *s) : p_(s) {}
    path(const std::filesystem::path &p) : p_(p) {}
    path parent_path() const { return path(p_.parent_path()); }
    path filename() const { return path(p_.filename()); }
    path stem() const { return path(p_.stem()); }
    bool is_complete() const { return p_.is_absolute(); }
    std::string string() const { return p_.string(); }
    friend path operator/(const path &a, const path &b) { return path(a.p_ / b.p_); }
};
inline path initial_path() { return path(std::filesystem::current_path()); }
inline bool exists(const path &p) { return std::filesystem::exists(p.string()); }
}  // namespace filesystem
}  // namespace boost
#endif
