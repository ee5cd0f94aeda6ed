/* boost/filesystem.hpp — SHIM (string based, POSIX separators) for the build of the reference's parameters.cpp / frames.cpp:
 * Params::get_filepath, the default .tnx index name, exists() */
#ifndef ORACLE_SHIM_PARAMS_BOOST_FILESYSTEM_HPP
#define ORACLE_SHIM_PARAMS_BOOST_FILESYSTEM_HPP
#include <sys/stat.h>
#include <unistd.h>
#include <string>
namespace boost {
namespace filesystem {
class path {
    std::string p_;
   public:
    path() {}
    path(const std::string &s) : p_(s) {}
    path(const char *s) : p_(s) {}
    path parent_path() const {
        size_t k = p_.find_last_of('/');
        if (k == std::string::npos) return path();
        return path(k == 0 ? std::string("/") : p_.substr(0, k));
    }
    path filename() const {
        size_t k = p_.find_last_of('/');
        return path(k == std::string::npos ? p_ : p_.substr(k + 1));
    }
    path stem() const {
        std::string f = filename().string();
        size_t k = f.find_last_of('.');
        return path((k == std::string::npos || k == 0 || f == "..") ? f : f.substr(0, k));
    }
    bool is_complete() const { return !p_.empty() && p_[0] == '/'; }
    bool empty() const { return p_.empty(); }
    std::string string() const { return p_; }
    friend path operator/(const path &a, const path &b) {
        if (b.is_complete() || a.p_.empty()) return b;
        if (b.p_.empty()) return a;
        return path(a.p_[a.p_.size() - 1] == '/' ? a.p_ + b.p_ : a.p_ + "/" + b.p_);
    }
};
inline path initial_path() {
    char buf[4096];
    return path(getcwd(buf, sizeof buf) ? buf : ".");
}
inline bool exists(const path &p) {
    struct stat st;
    return stat(p.string().c_str(), &st) == 0;
}
}  // namespace filesystem
}  // namespace boost
#endif
