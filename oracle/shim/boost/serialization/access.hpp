/* boost/serialization/access.hpp — SHIM: the reference's headers befriend this class and name base_object<> inside serialize()
 * templates that are never instantiated in the code built here. */
#ifndef ORACLE_SHIM_BOOST_SERIALIZATION_ACCESS_HPP
#define ORACLE_SHIM_BOOST_SERIALIZATION_ACCESS_HPP
namespace boost { namespace serialization {
class access;
template <class Base, class Derived>
Base &base_object(Derived &d) { return static_cast<Base &>(d); }
}}
#endif
