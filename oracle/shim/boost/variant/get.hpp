// oracle shim: see variant.hpp
#include <boost/variant/variant.hpp>
