/* boost/asio.hpp — SHIM: the scatter devices only pass tcp endpoints through to the (shimmed) service clients */
#ifndef ORACLE_SHIM_BOOST_ASIO_HPP
#define ORACLE_SHIM_BOOST_ASIO_HPP
namespace boost { namespace asio { namespace ip { namespace tcp { class endpoint {}; } } } }
#endif
