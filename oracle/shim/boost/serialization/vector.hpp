/* boost/serialization/vector.hpp — empty SHIM: the reference's serialize() templates are never instantiated in the code built here */
