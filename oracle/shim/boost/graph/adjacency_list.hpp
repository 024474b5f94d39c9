// oracle/shim/boost/graph/adjacency_list.hpp - the slice of the Boost Graph Library that the reference's topologies use
// (include/pagmo/topology.hpp:109-117: adjacency_list<vecS, vecS, bidirectionalS, no_property, double, no_property, listS>;
// src/topologies/base_bgl_topology.cpp, fully_connected.cpp:117-136, free_form.cpp), so that the UNMODIFIED topology.cpp,
// base_bgl_topology.cpp, ring.cpp, ... compile in this Boost-less image.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
// Semantics kept: vertices are 0..n-1; out-edges / in-edges of a vertex are iterated in insertion order (vecS), parallel
// edges are allowed by add_edge (base_bgl_topology refuses them itself), edge descriptors stay valid until the edge is
// removed (listS edge list), g[e] is the double edge property.
#ifndef ORACLE_SHIM_BOOST_GRAPH_ADJACENCY_LIST_HPP
#define ORACLE_SHIM_BOOST_GRAPH_ADJACENCY_LIST_HPP
#include <algorithm>
#include <cstddef>
#include <list>
#include <utility>
#include <vector>

namespace boost
{
struct vecS {};
struct listS {};
struct bidirectionalS {};
struct no_property {};

template <typename OutEdgeList, typename VertexList, typename Directed, typename VP, typename EP, typename GP, typename EdgeList>
class adjacency_list
{
    struct edge_node {
        std::size_t src, dst;
        EP prop;
    };
    using edge_list = std::list<edge_node>;

public:
    using vertex_descriptor = std::size_t;
    using vertices_size_type = std::size_t;
    using edges_size_type = std::size_t;
    struct edge_descriptor {
        std::size_t src = 0, dst = 0;
        edge_node *node = nullptr;
        bool operator==(const edge_descriptor &o) const { return node == o.node; }
        bool operator!=(const edge_descriptor &o) const { return node != o.node; }
    };

    adjacency_list() = default;
    explicit adjacency_list(std::size_t n) : m_out(n), m_in(n) {}
    adjacency_list(const adjacency_list &o) { copy_from(o); }
    adjacency_list(adjacency_list &&) noexcept = default;
    adjacency_list &operator=(const adjacency_list &o)
    {
        if (this != &o) {
            m_edges.clear();
            copy_from(o);
        }
        return *this;
    }
    adjacency_list &operator=(adjacency_list &&) noexcept = default;

    EP &operator[](const edge_descriptor &e) { return e.node->prop; }
    const EP &operator[](const edge_descriptor &e) const { return e.node->prop; }

    // ---- iterators -------------------------------------------------------------------------------------
    class vertex_iterator
    {
    public:
        using iterator_category = std::forward_iterator_tag;
        using value_type = std::size_t;
        using difference_type = std::ptrdiff_t;
        using pointer = const std::size_t *;
        using reference = std::size_t;
        explicit vertex_iterator(std::size_t i = 0) : m_i(i) {}
        std::size_t operator*() const { return m_i; }
        vertex_iterator &operator++() { ++m_i; return *this; }
        bool operator==(const vertex_iterator &o) const { return m_i == o.m_i; }
        bool operator!=(const vertex_iterator &o) const { return m_i != o.m_i; }
    private:
        std::size_t m_i;
    };
    using stored = std::vector<typename edge_list::iterator>;
    // walks a vertex's out- (or in-) edge vector and yields edge descriptors, targets or sources
    template <int What> // 0: edge descriptor, 1: target vertex, 2: source vertex
    class incident_iterator
    {
    public:
        using iterator_category = std::forward_iterator_tag;
        using value_type = typename std::conditional<What == 0, edge_descriptor, std::size_t>::type;
        using difference_type = std::ptrdiff_t;
        using pointer = const value_type *;
        using reference = value_type;
        incident_iterator() = default;
        explicit incident_iterator(typename stored::const_iterator it) : m_it(it) {}
        value_type operator*() const { return get(std::integral_constant<int, What>{}); }
        incident_iterator &operator++() { ++m_it; return *this; }
        bool operator==(const incident_iterator &o) const { return m_it == o.m_it; }
        bool operator!=(const incident_iterator &o) const { return m_it != o.m_it; }
    private:
        edge_descriptor get(std::integral_constant<int, 0>) const { return edge_descriptor{(*m_it)->src, (*m_it)->dst, &**m_it}; }
        std::size_t get(std::integral_constant<int, 1>) const { return (*m_it)->dst; }
        std::size_t get(std::integral_constant<int, 2>) const { return (*m_it)->src; }
        typename stored::const_iterator m_it;
    };
    using out_edge_iterator = incident_iterator<0>;
    using adjacency_iterator = incident_iterator<1>;
    using inv_adjacency_iterator = incident_iterator<2>;
    class edge_iterator
    {
    public:
        using iterator_category = std::forward_iterator_tag;
        using value_type = edge_descriptor;
        using difference_type = std::ptrdiff_t;
        using pointer = const edge_descriptor *;
        using reference = edge_descriptor;
        edge_iterator() = default;
        explicit edge_iterator(typename edge_list::iterator it) : m_it(it) {}
        edge_descriptor operator*() const { return edge_descriptor{m_it->src, m_it->dst, &*m_it}; }
        edge_iterator &operator++() { ++m_it; return *this; }
        bool operator==(const edge_iterator &o) const { return m_it == o.m_it; }
        bool operator!=(const edge_iterator &o) const { return m_it != o.m_it; }
    private:
        typename edge_list::iterator m_it;
    };

    // ---- the free functions below are friends through these members --------------------------------------
    std::size_t nv() const { return m_out.size(); }
    std::size_t ne() const { return m_edges.size(); }
    std::size_t add_v()
    {
        m_out.emplace_back();
        m_in.emplace_back();
        return m_out.size() - 1u;
    }
    std::pair<edge_descriptor, bool> add_e(std::size_t u, std::size_t v)
    {
        m_edges.push_back(edge_node{u, v, EP{}});
        auto it = std::prev(m_edges.end());
        m_out[u].push_back(it);
        m_in[v].push_back(it);
        return {edge_descriptor{u, v, &*it}, true};
    }
    std::pair<edge_descriptor, bool> find_e(std::size_t u, std::size_t v) const
    {
        for (const auto &it : m_out[u])
            if (it->dst == v) return {edge_descriptor{u, v, &*it}, true};
        return {edge_descriptor{}, false};
    }
    void remove_e(std::size_t u, std::size_t v) // removes every u -> v edge, as boost::remove_edge(u, v, g) does
    {
        for (auto oit = m_out[u].begin(); oit != m_out[u].end();) {
            if ((*oit)->dst == v) {
                auto node = *oit;
                auto &in = m_in[v];
                in.erase(std::find(in.begin(), in.end(), node));
                oit = m_out[u].erase(oit);
                m_edges.erase(node);
            } else
                ++oit;
        }
    }
    const stored &out_of(std::size_t v) const { return m_out[v]; }
    const stored &in_of(std::size_t v) const { return m_in[v]; }
    edge_list &edge_storage() const { return m_edges; }

private:
    void copy_from(const adjacency_list &o)
    {
        m_out.assign(o.m_out.size(), stored{});
        m_in.assign(o.m_in.size(), stored{});
        // global edge order first, then the per-vertex orders rebuilt through an old-node -> new-node map
        std::vector<std::pair<const edge_node *, typename edge_list::iterator>> map;
        for (const auto &e : o.m_edges) {
            m_edges.push_back(e);
            map.emplace_back(&e, std::prev(m_edges.end()));
        }
        std::sort(map.begin(), map.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
        auto lookup = [&map](const edge_node *p) {
            return std::lower_bound(map.begin(), map.end(), p, [](const auto &a, const edge_node *q) { return a.first < q; })->second;
        };
        for (std::size_t v = 0; v < o.m_out.size(); ++v) {
            for (const auto &it : o.m_out[v]) m_out[v].push_back(lookup(&*it));
            for (const auto &it : o.m_in[v]) m_in[v].push_back(lookup(&*it));
        }
    }
    mutable edge_list m_edges;
    std::vector<stored> m_out, m_in;
};

#define ORACLE_BGL_T template <typename A, typename B, typename C, typename D, typename E, typename F, typename G>
#define ORACLE_BGL_G adjacency_list<A, B, C, D, E, F, G>

ORACLE_BGL_T inline std::size_t num_vertices(const ORACLE_BGL_G &g) { return g.nv(); }
ORACLE_BGL_T inline std::size_t num_edges(const ORACLE_BGL_G &g) { return g.ne(); }
ORACLE_BGL_T inline std::size_t add_vertex(ORACLE_BGL_G &g) { return g.add_v(); }
ORACLE_BGL_T inline std::size_t vertex(std::size_t i, const ORACLE_BGL_G &) { return i; }
ORACLE_BGL_T inline std::pair<typename ORACLE_BGL_G::edge_descriptor, bool> add_edge(std::size_t u, std::size_t v, ORACLE_BGL_G &g)
{
    return g.add_e(u, v);
}
ORACLE_BGL_T inline std::pair<typename ORACLE_BGL_G::edge_descriptor, bool> edge(std::size_t u, std::size_t v, const ORACLE_BGL_G &g)
{
    return g.find_e(u, v);
}
ORACLE_BGL_T inline void remove_edge(std::size_t u, std::size_t v, ORACLE_BGL_G &g) { g.remove_e(u, v); }
ORACLE_BGL_T inline std::pair<typename ORACLE_BGL_G::vertex_iterator, typename ORACLE_BGL_G::vertex_iterator> vertices(const ORACLE_BGL_G &g)
{
    using it = typename ORACLE_BGL_G::vertex_iterator;
    return {it(0), it(g.nv())};
}
ORACLE_BGL_T inline std::pair<typename ORACLE_BGL_G::edge_iterator, typename ORACLE_BGL_G::edge_iterator> edges(const ORACLE_BGL_G &g)
{
    using it = typename ORACLE_BGL_G::edge_iterator;
    return {it(g.edge_storage().begin()), it(g.edge_storage().end())};
}
ORACLE_BGL_T inline std::pair<typename ORACLE_BGL_G::out_edge_iterator, typename ORACLE_BGL_G::out_edge_iterator> out_edges(std::size_t v,
                                                                                                                         const ORACLE_BGL_G &g)
{
    using it = typename ORACLE_BGL_G::out_edge_iterator;
    return {it(g.out_of(v).begin()), it(g.out_of(v).end())};
}
ORACLE_BGL_T inline std::pair<typename ORACLE_BGL_G::adjacency_iterator, typename ORACLE_BGL_G::adjacency_iterator>
adjacent_vertices(std::size_t v, const ORACLE_BGL_G &g)
{
    using it = typename ORACLE_BGL_G::adjacency_iterator;
    return {it(g.out_of(v).begin()), it(g.out_of(v).end())};
}
ORACLE_BGL_T inline std::pair<typename ORACLE_BGL_G::inv_adjacency_iterator, typename ORACLE_BGL_G::inv_adjacency_iterator>
inv_adjacent_vertices(std::size_t v, const ORACLE_BGL_G &g)
{
    using it = typename ORACLE_BGL_G::inv_adjacency_iterator;
    return {it(g.in_of(v).begin()), it(g.in_of(v).end())};
}
ORACLE_BGL_T inline std::size_t target(const typename ORACLE_BGL_G::edge_descriptor &e, const ORACLE_BGL_G &) { return e.dst; }
ORACLE_BGL_T inline std::size_t source(const typename ORACLE_BGL_G::edge_descriptor &e, const ORACLE_BGL_G &) { return e.src; }

#undef ORACLE_BGL_T
#undef ORACLE_BGL_G
} // namespace boost
#endif
