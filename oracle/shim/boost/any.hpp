// oracle shim: boost::any mapped onto std::any (island.hpp:46, not_population_based.hpp:35).
#ifndef ORACLE_SHIM_BOOST_ANY_HPP
#define ORACLE_SHIM_BOOST_ANY_HPP
#include <any>
namespace boost {
using any = std::any;
using bad_any_cast = std::bad_any_cast;
using std::any_cast;
}
#endif
