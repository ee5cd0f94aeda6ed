/* boost/serialization/map.hpp — empty SHIM: the reference's serialize() templates are never instantiated in the code built here */
