// oracle shim: boost::variant is absent in this image; the reference's migration policies (detail/base_sr_policy.hpp:36,
// fair_replace.cpp:83-90) only use construction/assignment, which() and boost::get<T>() - provided here over std::variant.
#ifndef ORACLE_SHIM_BOOST_VARIANT_HPP
#define ORACLE_SHIM_BOOST_VARIANT_HPP
#include <variant>
namespace boost {
template <typename... Ts> class variant : public std::variant<Ts...> {
public:
    using std::variant<Ts...>::variant;
    using std::variant<Ts...>::operator=;
    int which() const { return static_cast<int>(this->index()); }
};
template <typename T, typename... Ts> const T &get(const variant<Ts...> &v) { return std::get<T>(static_cast<const std::variant<Ts...> &>(v)); }
template <typename T, typename... Ts> T &get(variant<Ts...> &v) { return std::get<T>(static_cast<std::variant<Ts...> &>(v)); }
}
#endif
