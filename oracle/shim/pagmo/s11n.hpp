// oracle shim: Boost.Serialization is absent in this image; no archive types exist, serialize() templates stay uninstantiated.
// Replaces reference include/pagmo/s11n.hpp; every UDP declares
// `friend class boost::serialization::access` and a `serialize()` template (e.g. rastrigin.cpp:160-164).
#ifndef PAGMO_S11N_HPP
#define PAGMO_S11N_HPP
#include <utility>
namespace boost { namespace serialization {
class access {};
template <class B, class D> inline B &base_object(D &d) { return d; }
}}
// detail::archive / to_archive / from_archive: the reference's own Boost-free wrappers (`ar & x`, `ar << x`, `ar >> x` per
// argument).  No serialize() template is ever instantiated by the reference sources under this shim (the export macros
// below are no-ops); tests/cpp drives them with a small in-memory archive to round-trip the pagmo_cuda adapters.
#include <pagmo/detail/s11n_wrappers.hpp>
// Export / tracking macros of Boost.Serialization used by the real problem.hpp/bfe.hpp/algorithm.hpp
// (problem.hpp:67-75,886,1630): all no-ops here.
#define BOOST_CLASS_EXPORT_KEY2(T, K)
#define BOOST_CLASS_EXPORT_KEY(T)
#define BOOST_CLASS_EXPORT_IMPLEMENT(T)
#define BOOST_CLASS_TRACKING(T, E)
#define BOOST_CLASS_VERSION(T, N)
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
#define BOOST_SERIALIZATION_SPLIT_FREE(T)
#define BOOST_SERIALIZATION_ASSUME_ABSTRACT(T)
#endif
