/* io/xml_interface.hpp — SHIM.  The reference wraps libxml2 XPath (absent here).  Database::read_xml is compiled but never
 * called in oracle/_ref: the test harness registers the tables through the reference's own reg() methods, so this stand-in only
 * has to satisfy the compiler. */
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
    std::vector<XMLElement> get(const char *) { return std::vector<XMLElement>(); }
    void set_current(XMLElement) {}
    template <class T> T get_value(const char *) { return T(); }
};
#endif
