/* boost/serialization/access.hpp — SHIM: the reference's math/coor3d.hpp only befriends this class (its serialize()
 * templates are never instantiated here). */
#ifndef ORACLE_SHIM_BOOST_SERIALIZATION_ACCESS_HPP
#define ORACLE_SHIM_BOOST_SERIALIZATION_ACCESS_HPP
namespace boost { namespace serialization { class access; } }
#endif
