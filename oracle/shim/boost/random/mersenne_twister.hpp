/* boost/random/mersenne_twister.hpp — SHIM: boost::mt19937 is the standard MT19937 (same parameters, same seeding) */
#ifndef ORACLE_SHIM_BOOST_MT_HPP
#define ORACLE_SHIM_BOOST_MT_HPP
#include <random>
namespace boost { typedef std::mt19937 mt19937; }
#endif
