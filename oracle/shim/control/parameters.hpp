/* control/parameters.hpp — SHIM: the plain-struct Params of oracle/shim/control.hpp (+ lexical_cast, which the reference's
 * header brings in) */
#include <boost/lexical_cast.hpp>
#include "control.hpp"
