// oracle shim: stands in for the CMake-generated pagmo/config.hpp (reference config.hpp.in:36-47).
// TEST INFRASTRUCTURE ONLY - lets unmodified reference sources compile without Boost/TBB/Eigen.
#ifndef PAGMO_CONFIG_HPP
#define PAGMO_CONFIG_HPP
#define PAGMO_VERSION_STRING "2.19.1"
#define PAGMO_VERSION_MAJOR 2
#define PAGMO_VERSION_MINOR 19
#define PAGMO_VERSION_PATCH 1
#endif
