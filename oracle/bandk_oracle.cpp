// ============================================================================
// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's Band-k / CSR-k builder.
//
// Nothing in the product links, loads or calls this file; it is the checker used by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
//
// Parity status: PINNED against the reference itself (oracle/_ref/libref_csrk_*.so, built
// from the unmodified /root/reference/spmv-csrk/csrk.cpp) in tests/test_oracle_vs_ref.py,
// and against tests/golden/ (SURVEY.md Appendix B known-answer CRCs + generated vectors).
//
// Why C++ and not C: the RCM neighbour order (csrk.cpp:2345-2346) is whatever libstdc++'s
// non-stable std::sort does with the comparator `a.deg > b.deg` on {deg,id} pairs.  The only
// faithful restatement of "tie order of std::sort" is to call std::sort on the identical
// array with the identical comparator and element layout.
//
// What is restated (file:line relative to /root/reference/spmv-csrk/csrk.cpp):
//   threshold()          :896-897   T_i = (int)(sizes[i-1] * NNZ_{i-1} / N_{i-1})
//   group_rows()         :1255-1292 greedy contiguous grouping  (handCoarsen, first half)
//   coarse_graph()       :1295-1420 upper-edge symmetric closure, per-vertex sort, RLE dedup
//   level_structure()    :2426-2475 findRootedLevelStructures
//   pseudo_peripheral()  :2377-2421 findPseudoPeripheralVertex (George-Liu)
//   rcm_component()      :2289-2374 rcm_reordering_g
//   renumber()           :3012-3115 renumberGraphUsingReorderedVertices
//   uncoarsen()          :1148-1224 uncoarsenTheGraph
//   bandk_oracle_build   :841-1067  BAND_k::preprocessingForSpMV (HAND coarsening only)
//   bandk_oracle_reorder :548-676   CSRk_Graph::reorderA
//
// Deliberate difference: where the reference reads out of bounds (a coarse vertex with an
// empty adjacency list, csrk.cpp:1381-1385 -- only possible for inputs without a stored
// diagonal/upper entry) this returns error 2 instead.
// ============================================================================
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

typedef unsigned int u32;

struct Graph {
  long n = 0;
  std::vector<u32> rp;   // n+1
  std::vector<u32> adj;  // neighbours
  std::vector<u32> mult; // edge multiplicity ("degree" in the reference)
  long nnz() const { return (long)adj.size(); }
};

// csrk.cpp:896-897 -- unsigned * long / long, truncated to int
int threshold(u32 size, long nnz, long n) { return (int)((long)size * nnz / n); }

// csrk.cpp:1255-1292.  start has one entry per group plus the closing N.
void group_rows(const Graph &g, int T, std::vector<u32> &start, std::vector<u32> &grp) {
  grp.assign((size_t)g.n, 0);
  start.assign(1, 0);
  u32 acc = 0, cur = 0;
  for (long v = 0; v < g.n; ++v) {
    u32 len = g.rp[v + 1] - g.rp[v];
    if (acc < (u32)T) {
      acc += len;
    } else {
      ++cur;
      acc = len;
      start.push_back((u32)v);
    }
    grp[v] = cur;
  }
  if (acc > 0) start.push_back((u32)g.n);
  // (when acc == 0 at the end the reference leaves the last group unclosed; mirror it)
}

// csrk.cpp:1295-1420
int coarse_graph(const Graph &fine, const std::vector<u32> &start, const std::vector<u32> &grp,
                 Graph &out) {
  long nc = (long)start.size() - 1;
  std::vector<std::vector<u32>> lists((size_t)nc);
  for (long cv = 0; cv < nc; ++cv)
    for (u32 u = start[cv]; u < start[cv + 1]; ++u)
      for (u32 j = fine.rp[u]; j < fine.rp[u + 1]; ++j) {
        u32 w = fine.adj[j];
        if (w < start[cv]) continue; // edge into an earlier group: taken from the other side
        u32 t = grp[w];
        lists[cv].push_back(t);
        if ((long)t != cv) lists[t].push_back((u32)cv);
      }
  out.n = nc;
  out.rp.assign((size_t)nc + 1, 0);
  out.adj.clear();
  out.mult.clear();
  for (long cv = 0; cv < nc; ++cv) {
    std::vector<u32> &l = lists[cv];
    if (l.empty()) return 2;
    std::sort(l.begin(), l.end());
    size_t i = 0;
    while (i < l.size()) {
      size_t j = i;
      while (j < l.size() && l[j] == l[i]) ++j;
      out.adj.push_back(l[i]);
      out.mult.push_back((u32)(j - i));
      i = j;
    }
    out.rp[cv + 1] = (u32)out.adj.size();
  }
  return 0;
}

// csrk.cpp:2426-2475.  Returns the number of levels; lvl_ptr[0..levels], order = BFS order.
int level_structure(int root, const Graph &g, std::vector<int> &mask, std::vector<int> &lvl_ptr,
                    std::vector<int> &order) {
  order.clear();
  lvl_ptr.clear();
  mask[root] = 0;
  order.push_back(root);
  int levels = 0;
  size_t begin = 0, end = 0;
  long width = 1;
  while (width > 0) {
    begin = end;
    lvl_ptr.push_back((int)begin);
    end = order.size();
    for (size_t i = begin; i < end; ++i) {
      int v = order[i];
      for (u32 j = g.rp[v]; j < g.rp[v + 1]; ++j) {
        int w = (int)g.adj[j];
        if (mask[w] != 0) {
          order.push_back(w);
          mask[w] = 0;
        }
      }
    }
    width = (long)order.size() - (long)end;
    ++levels;
  }
  lvl_ptr.push_back((int)end);
  for (size_t i = 0; i < order.size(); ++i) mask[order[i]] = 1;
  return levels;
}

// csrk.cpp:2377-2421
void pseudo_peripheral(int &root, const Graph &g, std::vector<int> &mask) {
  std::vector<int> lvl_ptr, order;
  int levels = level_structure(root, g, mask, lvl_ptr, order);
  int cc = lvl_ptr[levels];
  if (levels == 1 || levels == cc) return;
  for (;;) {
    int j1 = lvl_ptr[levels - 1];
    int min_deg = cc;
    root = order[j1];
    if (cc != j1) {
      for (int j = j1; j < cc; ++j) {
        int v = order[j];
        int d = 0;
        for (u32 e = g.rp[v]; e < g.rp[v + 1]; ++e)
          if (mask[g.adj[e]] > 0) ++d;
        if (d < min_deg) {
          root = v;
          min_deg = d;
        }
      }
    }
    int new_levels = level_structure(root, g, mask, lvl_ptr, order);
    if (new_levels <= levels || levels >= cc) return;
    levels = new_levels;
  }
}

// element type and comparator exactly as csrk.h:477-490 / csrk.cpp:65-67
struct DegId {
  unsigned deg;
  unsigned id;
};
bool by_deg_desc(DegId a, DegId b) { return a.deg > b.deg; }

// csrk.cpp:2289-2374
void rcm_component(const Graph &g, int root, std::vector<int> &mask, std::vector<u32> &old2new,
                   std::vector<u32> &new2old, int &done) {
  pseudo_peripheral(root, g, mask);
  std::vector<int> queue;
  queue.push_back(root);
  size_t head = 0;
  int visited = 0;
  int pos = done;
  while (head < queue.size()) {
    int p = queue[head++];
    if (mask[p] != 1) continue;
    mask[p] = 0;
    new2old[pos++] = (u32)p;
    ++visited;
    std::vector<DegId> cand(g.rp[p + 1] - g.rp[p]);
    int cnt = 0;
    for (u32 e = g.rp[p]; e < g.rp[p + 1]; ++e) {
      int q = (int)g.adj[e];
      if (mask[q] == 1) {
        cand[cnt].deg = g.mult[e];
        cand[cnt].id = (unsigned)q;
        ++cnt;
      }
    }
    std::sort(cand.begin(), cand.begin() + cnt, by_deg_desc);
    for (int i = 0; i < cnt; ++i) queue.push_back((int)cand[i].id);
  }
  // reverse the segment [done, done+visited) and fill old->new
  int half = visited / 2, last = pos - 1;
  for (int i = 0; i < half; ++i) {
    std::swap(new2old[last], new2old[done + i]);
    old2new[new2old[last]] = (u32)last;
    old2new[new2old[done + i]] = (u32)(done + i);
    --last;
  }
  if (visited % 2 == 1) old2new[new2old[done + half]] = (u32)(done + half);
  done += visited;
}

// csrk.cpp:917-933: RCM over every connected component in increasing root order
void rcm_all(const Graph &g, std::vector<u32> &old2new, std::vector<u32> &new2old) {
  std::vector<int> mask((size_t)g.n, 1);
  old2new.assign((size_t)g.n, 0);
  new2old.assign((size_t)g.n, 0);
  int done = 0;
  for (long v = 0; v < g.n; ++v) {
    if (done >= g.n) break;
    if (mask[v] != 0) rcm_component(g, (int)v, mask, old2new, new2old, done);
  }
}

// element type / comparator as csrk.h:57-70 / csrk.cpp:61 (`<=`; ids are distinct after dedup)
struct AdjDeg {
  int adj;
  int deg;
};
bool adj_le(AdjDeg a, AdjDeg b) { return a.adj <= b.adj; }

// csrk.cpp:3012-3115
void renumber(Graph &g, const std::vector<u32> &old2new, const std::vector<u32> &new2old) {
  Graph src = g;
  g.rp[0] = 0;
  for (long i = 0; i < g.n; ++i) {
    u32 o = new2old[i];
    g.rp[i + 1] = g.rp[i] + (src.rp[o + 1] - src.rp[o]);
  }
  for (long i = 0; i < g.n; ++i) {
    u32 o = new2old[i];
    std::vector<AdjDeg> tmp;
    for (u32 j = src.rp[o]; j < src.rp[o + 1]; ++j) {
      AdjDeg a;
      a.adj = (int)old2new[src.adj[j]];
      a.deg = (int)src.mult[j];
      tmp.push_back(a);
    }
    std::sort(tmp.begin(), tmp.end(), adj_le);
    for (size_t t = 0; t < tmp.size(); ++t) {
      g.adj[g.rp[i] + t] = (u32)tmp[t].adj;
      g.mult[g.rp[i] + t] = (u32)tmp[t].deg;
    }
  }
}

// csrk.cpp:1148-1224.  map: level -> level-1 group starts (rewritten in place);
// perm_coarse: new->old of the coarse level; perm_fine: new->old of the finer level (composed).
void uncoarsen(std::vector<u32> &map, const std::vector<u32> &perm_coarse,
               std::vector<u32> &perm_fine) {
  long nc = (long)perm_coarse.size();
  std::vector<u32> old_map = map;
  std::vector<u32> q(perm_fine.size());
  map[0] = 0;
  for (long j = 0; j < nc; ++j) {
    u32 o = perm_coarse[j];
    map[j + 1] = map[j] + (old_map[o + 1] - old_map[o]);
    u32 dst = map[j];
    for (u32 t = old_map[o]; t < old_map[o + 1]; ++t) q[dst++] = t;
  }
  std::vector<u32> old_perm = perm_fine;
  for (size_t i = 0; i < perm_fine.size(); ++i) perm_fine[i] = old_perm[q[i]];
}

} // namespace

extern "C" {

// Band-k for k in {2,3} with HAND coarsening.  Outputs:
//   perm[n]               new row -> old row           (permBigG)
//   map1[n1+1], *n1       super-row -> row starts      (mapCoarseToFinerRows[1])
//   map2[n2+1], *n2       super-super-row -> super-row (mapCoarseToFinerRows[2]; k == 3 only)
// map1/map2 must have room for n+1 entries.  Returns 0, or 1 (bad arguments), 2 (coarse
// vertex without adjacency: reference would read out of bounds).
int bandk_oracle_build(long n, long nnz, const u32 *r, const u32 *c, int k, const int *sizes,
                       u32 *perm, u32 *map1, long *n1, u32 *map2, long *n2) {
  if (k < 2 || k > 3 || n <= 0) return 1;
  std::vector<Graph> g((size_t)k);
  g[0].n = n;
  g[0].rp.assign(r, r + n + 1);
  g[0].adj.assign(c, c + nnz);
  g[0].mult.assign((size_t)nnz, 1);

  std::vector<std::vector<u32>> perms((size_t)k), maps((size_t)k);
  perms[0].resize((size_t)n);
  for (long i = 0; i < n; ++i) perms[0][i] = (u32)i;

  for (int lvl = 1; lvl < k; ++lvl) {
    const Graph &fine = g[lvl - 1];
    int T = threshold((u32)sizes[lvl - 1], fine.nnz(), fine.n);
    std::vector<u32> grp;
    group_rows(fine, T, maps[lvl], grp);
    int rc = coarse_graph(fine, maps[lvl], grp, g[lvl]);
    if (rc) return rc;
    std::vector<u32> o2n;
    rcm_all(g[lvl], o2n, perms[lvl]);
    renumber(g[lvl], o2n, perms[lvl]);
  }
  for (int lvl = k - 1; lvl >= 1; --lvl) uncoarsen(maps[lvl], perms[lvl], perms[lvl - 1]);

  std::memcpy(perm, perms[0].data(), sizeof(u32) * (size_t)n);
  *n1 = (long)maps[1].size() - 1;
  std::memcpy(map1, maps[1].data(), sizeof(u32) * maps[1].size());
  if (k == 3) {
    *n2 = (long)maps[2].size() - 1;
    std::memcpy(map2, maps[2].data(), sizeof(u32) * maps[2].size());
  } else if (n2) {
    *n2 = 0;
  }
  return 0;
}

#define DEFINE_REORDER(T, SFX)                                                              \
  /* csrk.cpp:548-676: A' = P A P^T, each row sorted by new column id */                    \
  void bandk_oracle_reorder_##SFX(long n, const u32 *perm, const u32 *r, const u32 *c,      \
                                  const T *v, u32 *r2, u32 *c2, T *v2) {                    \
    std::vector<u32> fwd((size_t)n);                                                        \
    for (long i = 0; i < n; ++i) fwd[perm[i]] = (u32)i;                                     \
    r2[0] = 0;                                                                              \
    for (long i = 0; i < n; ++i) r2[i + 1] = r2[i] + (r[perm[i] + 1] - r[perm[i]]);         \
    for (long i = 0; i < n; ++i) {                                                          \
      u32 o = perm[i];                                                                      \
      std::vector<std::pair<int, u32>> row; /* (new col, old position) */                  \
      for (u32 j = r[o]; j < r[o + 1]; ++j) row.push_back(std::make_pair((int)fwd[c[j]], j)); \
      std::sort(row.begin(), row.end(),                                                     \
                [](const std::pair<int, u32> &a, const std::pair<int, u32> &b) {            \
                  return a.first < b.first;                                                 \
                });                                                                         \
      for (size_t t = 0; t < row.size(); ++t) {                                             \
        c2[r2[i] + t] = (u32)row[t].first;                                                  \
        v2[r2[i] + t] = v[row[t].second];                                                   \
      }                                                                                     \
    }                                                                                       \
  }

DEFINE_REORDER(float, f32)
DEFINE_REORDER(double, f64)

} // extern "C"
