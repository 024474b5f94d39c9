// oracle shim: boost::optional mapped onto std::optional (nsga2.hpp:36, pso_gen.hpp:36, ...).
#ifndef ORACLE_SHIM_BOOST_OPTIONAL_HPP
#define ORACLE_SHIM_BOOST_OPTIONAL_HPP
#include <optional>
namespace boost {
template <typename T> using optional = std::optional<T>;
inline constexpr std::nullopt_t none = std::nullopt;
}
#endif
