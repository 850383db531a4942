// TEST INFRASTRUCTURE ONLY -- compile-only stand-in for the two Boost.Graph headers that
// the reference's csrk.h includes (spmv-csrk/csrk.h:20-21).  Boost is not installed in this
// image.  The only reference function that touches these types is BGL_ordering
// (spmv-csrk/csrk.cpp:2946-3009), reachable only from the STS/COLOR path, never from SpMV
// or the Band-k builder, so the stub aborts if it is ever executed.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <utility>
#include <vector>

namespace boost {

struct listS {};
struct vecS {};
struct undirectedS {};
enum vertex_index_t { vertex_index };

template <class OutEdgeList, class VertexList, class Directed>
struct adjacency_list {
  std::size_t nv;
  template <class EdgeIt>
  adjacency_list(EdgeIt, EdgeIt, std::size_t n) : nv(n) {}
};

template <class G> struct graph_traits {
  typedef std::size_t vertices_size_type;
};

struct stub_index_map {};

template <class G, class Tag> struct property_map {
  typedef stub_index_map const_type;
  typedef stub_index_map type;
};

template <class G> inline stub_index_map get(vertex_index_t, const G &) {
  return stub_index_map();
}

template <class Iter, class IndexMap> struct iterator_property_map {
  Iter base;
  iterator_property_map(Iter b, IndexMap) : base(b) {}
  template <class K> auto operator[](K k) const -> decltype(base[k]) {
    return base[k];
  }
};

template <class G> inline std::size_t num_vertices(const G &g) { return g.nv; }

template <class G, class ColorMap>
inline std::size_t sequential_vertex_coloring(const G &, ColorMap) {
  std::fprintf(stderr, "boost stub: sequential_vertex_coloring is not available "
                       "(STS/COLOR path is out of scope)\n");
  std::abort();
  return 0;
}

} // namespace boost
