/* io/xml_interface.hpp — SHIM for the build of the reference's parameters.cpp: libxml2 is absent, Params::read_xml is compiled
 * but never called (the test harness fills the parameter structs and calls the generators). */
#ifndef ORACLE_SHIM_IO_XML_INTERFACE_HPP
#define ORACLE_SHIM_IO_XML_INTERFACE_HPP
#include <string>
#include <vector>
class XMLElement {};
class XMLInterface {
   public:
    explicit XMLInterface(std::string) {}
    void dump(std::vector<char> &) {}
    bool exists(const char *) { return false; }
    bool exists(const std::string &) { return false; }
    std::vector<XMLElement> get(const char *) { return std::vector<XMLElement>(); }
    std::vector<XMLElement> get(const std::string &) { return std::vector<XMLElement>(); }
    void set_current(XMLElement) {}
    template <class T> T get_value(const char *) { return T(); }
    template <class T> T get_value(const std::string &) { return T(); }
};
#endif
