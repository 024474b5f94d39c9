// oracle shim for boost::numeric_cast: range-checked static_cast (generic.hpp:40, multi_objective.hpp:41).
#ifndef ORACLE_SHIM_BOOST_NUMERIC_CAST_HPP
#define ORACLE_SHIM_BOOST_NUMERIC_CAST_HPP
#include <limits>
#include <stdexcept>
#include <type_traits>
namespace boost { namespace numeric {
struct bad_numeric_cast : std::bad_cast { const char *what() const noexcept override { return "bad numeric conversion"; } };
struct positive_overflow : bad_numeric_cast {};
struct negative_overflow : bad_numeric_cast {};
}
template <typename T, typename U> inline T numeric_cast(U u)
{
    if constexpr (std::is_integral<T>::value && std::is_integral<U>::value) {
        if constexpr (std::is_signed<U>::value && !std::is_signed<T>::value) {
            if (u < 0) throw numeric::negative_overflow{};
            if (static_cast<typename std::make_unsigned<U>::type>(u) > std::numeric_limits<T>::max()) throw numeric::positive_overflow{};
        } else if constexpr (!std::is_signed<U>::value && std::is_signed<T>::value) {
            if (u > static_cast<typename std::make_unsigned<T>::type>(std::numeric_limits<T>::max())) throw numeric::positive_overflow{};
        } else {
            if (u > std::numeric_limits<T>::max()) throw numeric::positive_overflow{};
            if (u < std::numeric_limits<T>::lowest()) throw numeric::negative_overflow{};
        }
    } else if constexpr (std::is_integral<T>::value) {
        if (!(u >= static_cast<U>(std::numeric_limits<T>::lowest()) - 1 && u < static_cast<U>(std::numeric_limits<T>::max()) + 1))
            throw numeric::positive_overflow{};
    }
    return static_cast<T>(u);
}
namespace numeric { using boost::numeric_cast; }
}
#endif
