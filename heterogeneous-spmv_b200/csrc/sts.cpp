// Host index work of the sparse triangular solve on CSR-k, k = 2, level-set ordering:
// what BAND_k::preprocessingForSTS does for kernel "STS" / order "LS" (spmv-csrk/csrk.cpp:1522-1672 with
// find_levels :2704-2816).  Must be bit-exact against oracle/sts_oracle.cpp, which restates the reference literally.
//
// The reference finds the level of a vertex by deleting processed neighbours from a COPY of the adjacency and
// re-scanning it (O(nnz * degree)); here a counter of unprocessed lower neighbours per vertex gives the same
// decisions in O(nnz log degree): "the first entry that is not deleted is the vertex itself" <=> "no lower
// neighbour is left" (columns are sorted, the diagonal is stored).  What has to be reproduced exactly is the ORDER
// inside a level -- the order in which candidates are discovered: level vertices in order, their neighbours in
// storage order, a vertex enters the candidate list once until it is examined and fails -- and the order of the
// levels: std::sort by size with the reference's comparator on the reference's element type (non-stable: the tie
// order is libstdc++'s, so the identical call is made).
#include "host_index.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace hspmv {

namespace {
struct degree_id_pair { // spmv-csrk/csrk.h:84-98
  int deg;
  int id;
  degree_id_pair() : deg(0), id(0) {}
  degree_id_pair(int a, int b) : deg(a), id(b) {}
};
bool compare_deg_id_pair(degree_id_pair p1, degree_id_pair p2) { return p1.deg < p2.deg; } // csrk.cpp:57-59
} // namespace

// perm: new -> old; pack_ptr: n_packs + 1 entries.  HSPMV_EUNSUPPORTED when a row has no stored diagonal or the
// search stalls (structurally unsymmetric matrix) -- the reference prints "Number of vertices did not change in
// level" and goes on with garbage (csrk.cpp:2797-2800).
// levels sorted by size with the reference's call, then written out as perm / pack_ptr (shared by both finders)
static void emit_packs(int64_t n, const std::vector<int32_t> &order, const std::vector<int32_t> &level_ptr, uvec &perm,
                       std::vector<int32_t> &pack_ptr) {
  const int n_levels = (int)level_ptr.size() - 1;
  std::vector<degree_id_pair> sizes((size_t)n_levels);
  for (int i = 0; i < n_levels; ++i) sizes[(size_t)i] = degree_id_pair(level_ptr[(size_t)i + 1] - level_ptr[(size_t)i], i);
  std::sort(sizes.begin(), sizes.begin() + n_levels, compare_deg_id_pair);
  perm.resize((size_t)n);
  pack_ptr.assign(1, 0);
  for (int i = 0; i < n_levels; ++i) {
    const int old = sizes[(size_t)i].id;
    pack_ptr.push_back(pack_ptr.back() + (level_ptr[(size_t)old + 1] - level_ptr[(size_t)old]));
  }
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < n_levels; ++i) {
    const int old = sizes[(size_t)i].id;
    int32_t dst = pack_ptr[(size_t)i];
    for (int32_t j = level_ptr[(size_t)old]; j < level_ptr[(size_t)old + 1]; ++j) perm[(size_t)dst++] = (u32)order[(size_t)j];
  }
}

// The same level sets for a STRUCTURALLY SYMMETRIC matrix, without the level-by-level bookkeeping.  With (u, v) stored
// whenever (v, u) is, the search above reduces to:
//   * level(v) = 0 when row v has nothing before its diagonal, else 1 + max level(u) over its lower neighbours u < v --
//     one pass over the rows in index order (a vertex enters the level after the one that removed its last lower
//     neighbour, and it is always a candidate then: that neighbour lists it);
//   * order inside level L+1 = the order in which the scan of level L (its vertices in order, their neighbours in
//     storage order) first meets each vertex of level L+1 -- candidates that fail are forgotten at the end of a level,
//     so only this scan decides; level 0 is in index order.
// The symmetry test (every entry's mirror found by binary search) runs on all cores; the two passes are O(nnz) with no
// searches.  Returns 1 when the matrix is not symmetric (the caller then runs the literal search, which also covers
// the inputs the reference accepts by accident), HSPMV_EUNSUPPORTED for a missing diagonal.
static int sts_levels_symmetric(int64_t n, const int32_t *rp, const int32_t *ci, std::vector<int32_t> &order,
                                std::vector<int32_t> &level_ptr) {
  std::vector<int32_t> diag((size_t)n);
  int bad_diag = 0, asym = 0;
#pragma omp parallel for schedule(dynamic, 4096) reduction(| : bad_diag, asym)
  for (int64_t v = 0; v < n; ++v) {
    const int32_t *b = ci + rp[v], *e = ci + rp[v + 1];
    const int32_t *d = std::lower_bound(b, e, (int32_t)v);
    if (d == e || *d != (int32_t)v) {
      bad_diag = 1;
      diag[(size_t)v] = 0;
      continue;
    }
    diag[(size_t)v] = (int32_t)(d - b);
    for (const int32_t *p = b; p < e; ++p) {
      const int32_t u = *p;
      if (u == (int32_t)v) continue;
      if (u < 0 || u >= n || (p > b && p[-1] >= u) || !std::binary_search(ci + rp[u], ci + rp[u + 1], (int32_t)v)) asym = 1;
    }
  }
  if (bad_diag) return HSPMV_EUNSUPPORTED;
  if (asym) return 1;
  std::vector<int32_t> level((size_t)n);
  int32_t n_levels = 0;
  for (int64_t v = 0; v < n; ++v) {
    int32_t lv = -1;
    const int32_t *b = ci + rp[v];
    for (int32_t t = 0; t < diag[(size_t)v]; ++t) lv = std::max(lv, level[(size_t)b[t]]);
    level[(size_t)v] = lv + 1;
    n_levels = std::max(n_levels, lv + 2);
  }
  level_ptr.assign((size_t)n_levels + 1, 0);
  for (int64_t v = 0; v < n; ++v) ++level_ptr[(size_t)level[(size_t)v] + 1];
  for (int32_t l = 0; l < n_levels; ++l) level_ptr[(size_t)l + 1] += level_ptr[(size_t)l];
  order.resize((size_t)n);
  int32_t filled = 0;
  for (int64_t v = 0; v < n; ++v)
    if (level[(size_t)v] == 0) order[(size_t)filled++] = (int32_t)v;
  std::vector<char> placed((size_t)n, 0);
  for (int32_t l = 0; l + 1 < n_levels; ++l) {
    int32_t dst = level_ptr[(size_t)l + 1];
    for (int32_t i = level_ptr[(size_t)l]; i < level_ptr[(size_t)l + 1]; ++i) {
      const int32_t g = order[(size_t)i];
      for (int32_t j = rp[g]; j < rp[g + 1]; ++j) {
        const int32_t nb = ci[j];
        if (level[(size_t)nb] == l + 1 && !placed[(size_t)nb]) {
          placed[(size_t)nb] = 1;
          order[(size_t)dst++] = nb;
        }
      }
    }
    if (dst != level_ptr[(size_t)l + 2]) return 1; // (cannot happen for a symmetric matrix; let the literal search decide)
  }
  return HSPMV_OK;
}

int sts_levels(int64_t n, const int32_t *rp, const int32_t *ci, uvec &perm, std::vector<int32_t> &pack_ptr) {
  if (n <= 0 || !rp || !ci) return HSPMV_EINVAL;
  const char *mode = std::getenv("HSPMV_STS_LEVELS"); // "literal" forces the level-by-level search (tests)
  if (!(mode && std::strcmp(mode, "literal") == 0)) {
    std::vector<int32_t> order, level_ptr;
    const int rc = sts_levels_symmetric(n, rp, ci, order, level_ptr);
    if (rc == HSPMV_OK) {
      emit_packs(n, order, level_ptr, perm, pack_ptr);
      return HSPMV_OK;
    }
    if (rc != 1) return rc;
  }
  std::vector<int32_t> lower((size_t)n);
  for (int64_t v = 0; v < n; ++v) { // entries before the diagonal; the diagonal must be there
    const int32_t *b = ci + rp[v], *e = ci + rp[v + 1];
    const int32_t *d = std::lower_bound(b, e, (int32_t)v);
    if (d == e || *d != (int32_t)v) return HSPMV_EUNSUPPORTED;
    lower[(size_t)v] = (int32_t)(d - b);
  }
  std::vector<char> processed((size_t)n, 0), not_candidate((size_t)n, 0);
  std::vector<int32_t> order, cand, next_cand, level_ptr(1, 0);
  order.reserve((size_t)n);
  for (int64_t v = 0; v < n; ++v)
    if (lower[(size_t)v] == 0) {
      order.push_back((int32_t)v);
      processed[(size_t)v] = not_candidate[(size_t)v] = 1;
    }
  level_ptr.push_back((int32_t)order.size());
  while ((int64_t)order.size() < n) {
    if (level_ptr[level_ptr.size() - 1] == level_ptr[level_ptr.size() - 2]) return HSPMV_EUNSUPPORTED; // stalled
    cand.clear();
    for (int32_t i = level_ptr[level_ptr.size() - 2]; i < level_ptr.back(); ++i) {
      const int32_t g = order[(size_t)i];
      for (int32_t j = rp[g]; j < rp[g + 1]; ++j) {
        const int32_t nb = ci[j];
        if (processed[(size_t)nb]) continue;
        if (!not_candidate[(size_t)nb]) {
          cand.push_back(nb);
          not_candidate[(size_t)nb] = 1;
        }
      }
    }
    // deletions of this level: every unprocessed neighbour nb > g that stores g loses one lower neighbour
    for (int32_t i = level_ptr[level_ptr.size() - 2]; i < level_ptr.back(); ++i) {
      const int32_t g = order[(size_t)i];
      for (int32_t j = rp[g]; j < rp[g + 1]; ++j) {
        const int32_t nb = ci[j];
        if (nb <= g || processed[(size_t)nb]) continue;
        const int32_t *b = ci + rp[nb], *e = ci + rp[nb + 1];
        if (std::binary_search(b, e, g)) --lower[(size_t)nb];
      }
    }
    for (int32_t v : cand) {
      if (lower[(size_t)v] == 0) {
        order.push_back(v);
        processed[(size_t)v] = 1;
      } else {
        not_candidate[(size_t)v] = 0;
      }
    }
    level_ptr.push_back((int32_t)order.size());
  }
  emit_packs(n, order, level_ptr, perm, pack_ptr);
  return HSPMV_OK;
}

} // namespace hspmv
