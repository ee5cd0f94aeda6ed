/* boost/thread/barrier.hpp — SHIM: see ../thread.hpp */
#include <boost/thread.hpp>
