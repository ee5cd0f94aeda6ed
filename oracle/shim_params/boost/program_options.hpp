/* boost/program_options.hpp — SHIM for the build of the reference's parameters.cpp (oracle/_ref/libparams_ref.so): enough of
 * the interface for Params::options / overwrite_options / init to compile.  Command lines are not parsed (the variables_map
 * stays empty); the test harness calls the generators directly. */
#ifndef ORACLE_SHIM_BOOST_PROGRAM_OPTIONS_HPP
#define ORACLE_SHIM_BOOST_PROGRAM_OPTIONS_HPP
#include <iostream>
#include <string>
#include <boost/lexical_cast.hpp>
namespace boost {
namespace program_options {
template <class T> struct typed_value {
    typed_value *default_value(const T &) { return this; }
};
template <class T> typed_value<T> *value() {
    static typed_value<T> v;
    return &v;
}
struct options_description;
struct options_description_easy_init {
    options_description_easy_init &operator()(const char *, const char *) { return *this; }
    template <class T> options_description_easy_init &operator()(const char *, typed_value<T> *, const char *) { return *this; }
};
struct options_description {
    options_description() {}
    explicit options_description(const std::string &) {}
    options_description_easy_init add_options() { return options_description_easy_init(); }
    options_description &add(const options_description &) { return *this; }
};
inline std::ostream &operator<<(std::ostream &os, const options_description &) { return os; }
struct variable_value {
    template <class T> T as() const { return T(); }
    bool defaulted() const { return true; }
};
struct variables_map {
    typedef const void *const_iterator;
    const_iterator find(const std::string &) const { return nullptr; }
    const_iterator end() const { return nullptr; }
    size_t count(const std::string &) const { return 0; }
    variable_value operator[](const std::string &) const { return variable_value(); }
};
struct parsed_options {};
inline parsed_options parse_command_line(int, char **, const options_description &) { return parsed_options(); }
inline void store(const parsed_options &, variables_map &) {}
inline void notify(variables_map &) {}
}  // namespace program_options
}  // namespace boost
#endif
