// oracle shim: Boost.Serialization support for adjacency_list - nothing to declare (no archives exist in this image).
#ifndef ORACLE_SHIM_BOOST_GRAPH_ADJ_LIST_SERIALIZE_HPP
#define ORACLE_SHIM_BOOST_GRAPH_ADJ_LIST_SERIALIZE_HPP
#include <boost/graph/adjacency_list.hpp>
#endif
