// Band-k / CSR-k builder (host side of hspmv_build_csrk, hspmv_bandk_host, hspmv_reorder_host).
//
// Produces, bit for bit, the index structures of the reference's HAND-coarsening path:
//   BAND_k::preprocessingForSpMV  spmv-csrk/csrk.cpp:841-1067
//   BAND_k::handCoarsen           :1243-1437      BAND_k::uncoarsenTheGraph :1148-1224
//   rcm_reordering_g              :2289-2374      findPseudoPeripheralVertex :2377-2421
//   findRootedLevelStructures     :2426-2475      renumberGraphUsingReorderedVertices :3012-3115
//   CSRk_Graph::reorderA          :548-676
// but is organised differently so that it scales to the 1.7 G-nnz configuration:
//   * the coarse graph is assembled without the reference's 2*nnz-entry scratch list and its
//     serial dedup: each coarse vertex's UPPER neighbours (same or later group) are collected,
//     sorted and run-length encoded in parallel, and the lower half is obtained by transposing
//     the strict upper half (the reference inserts the mirrored entry explicitly at
//     csrk.cpp:1350-1354; after its per-vertex sort the two constructions are identical:
//     entries < v come only from mirrors, entries >= v only from v's own rows);
//   * BFS bookkeeping uses epoch stamps instead of clearing and restoring a mask;
//   * reorderA sorts each row in a per-thread scratch buffer instead of an nnz-sized pair array.
// The one place where the result depends on library behaviour -- the order in which
// std::sort leaves equal-multiplicity neighbours (csrk.cpp:2345-2346, comparator :65-67) -- is
// reproduced by calling std::sort on the same {deg,id} element layout with the same predicate.
#include "host_index.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace hspmv {

namespace {

typedef CoarseGraph Coarse;

// HSPMV_BANDK_TRACE=1 prints the wall time of every phase of the build to stderr
struct Phase {
  const char *name;
  std::chrono::steady_clock::time_point t0;
  bool on;
  explicit Phase(const char *n) : name(n), t0(std::chrono::steady_clock::now()), on(std::getenv("HSPMV_BANDK_TRACE") != nullptr) {}
  ~Phase() {
    if (on) {
      std::chrono::duration<double> d = std::chrono::steady_clock::now() - t0;
      std::fprintf(stderr, "[bandk] %-28s %8.3f s\n", name, d.count());
    }
  }
};

// Greedy contiguous grouping (csrk.cpp:1278-1292): a group is closed by the vertex that makes
// the running nnz reach the threshold; the next vertex opens a new group.  This literal form is
// kept for T == 0 (every vertex, the first included, opens a group) and as the statement of the rule.
void group_rows_literal(const u32 *rp, int64_t n, u32 T, uvec &start, uvec &grp) {
  grp.resize((size_t)n);
  start.clear();
  start.push_back(0);
  u32 acc = 0, g = 0;
  for (int64_t v = 0; v < n; ++v) {
    const u32 len = rp[v + 1] - rp[v];
    if (acc < T) {
      acc += len;
    } else {
      ++g;
      acc = len;
      start.push_back((u32)v);
    }
    grp[(size_t)v] = g;
  }
  if (acc > 0) start.push_back((u32)n);
}

// The same group starts without the per-row bookkeeping: the running count of a group that starts at row s is
// rp[v] - rp[s], so (for T > 0) the next group starts at the first v > s with rp[v] >= rp[s] + T -- a forward scan of
// rp, 4x faster than the literal loop at 64 M rows (0.32 -> 0.08 s), which matters because the device builder
// (bandk_gpu.cu) only needs the starts.
void group_starts(const u32 *rp, int64_t n, u32 T, uvec &start) {
  if (T == 0 || n <= 0) {
    uvec grp;
    group_rows_literal(rp, n, T, start, grp);
    return;
  }
  start.clear();
  start.push_back(0);
  int64_t s = 0;
  for (;;) {
    const uint64_t target = (uint64_t)rp[s] + T;
    int64_t u = s + 1;
    while (u < n && (uint64_t)rp[u] < target) ++u;
    if (u >= n) break;
    start.push_back((u32)u);
    s = u;
  }
  if (rp[n] - rp[s] > 0) start.push_back((u32)n);
}

void group_rows(const u32 *rp, int64_t n, u32 T, uvec &start, uvec &grp) {
  if (T == 0 || n <= 0) {
    group_rows_literal(rp, n, T, start, grp);
    return;
  }
  group_starts(rp, n, T, start);
  grp.resize((size_t)n);
  const int64_t ng = (int64_t)start.size();
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < ng; ++j) {
    const int64_t hi = j + 1 < ng ? (int64_t)start[(size_t)j + 1] : n;
    for (int64_t v = start[(size_t)j]; v < hi; ++v) grp[(size_t)v] = (u32)j;
  }
}

// (rp, adj): the graph to coarsen; start / grp: the groups.  With new2old / old2new the graph is still stored in its
// OLD numbering while start / grp (and the result) are in the new one: vertex u is row new2old[u] and its neighbours
// are old2new[adj[..]] -- renumberGraphUsingReorderedVertices (csrk.cpp:3012-3115) applied on the fly.  The result
// does not depend on the order of a row's entries (they are collected, sorted and run-length encoded), so it equals
// coarsening the relabelled graph of relabel() below, without copying and re-sorting its 100 M entries at config 5.
int build_coarse(const u32 *rp, const u32 *adj, const uvec &start,
                 const uvec &grp, Coarse &out, const u32 *new2old = nullptr, const u32 *old2new = nullptr) {
  const int64_t nc = (int64_t)start.size() - 1;
  out.n = nc;
  uvec up_cnt((size_t)nc + 1, 0);
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  std::vector<uvec> buf_adj((size_t)nthreads), buf_mult((size_t)nthreads);
  std::vector<int64_t> first_cv((size_t)nthreads + 1, nc);

#pragma omp parallel num_threads(nthreads)
  {
    int tid = 0, nt = 1;
#ifdef _OPENMP
    tid = omp_get_thread_num();
    nt = omp_get_num_threads();
#endif
    const int64_t lo = nc * tid / nt, hi = nc * (tid + 1) / nt;
    first_cv[(size_t)tid] = lo;
    uvec scratch;
    uvec &oa = buf_adj[(size_t)tid], &om = buf_mult[(size_t)tid];
    for (int64_t cv = lo; cv < hi; ++cv) {
      scratch.clear();
      const u32 gs = start[(size_t)cv];
      if (new2old) {
        for (u32 u = gs; u < start[(size_t)cv + 1]; ++u) {
          const u32 o = new2old[u];
          for (u32 j = rp[o]; j < rp[o + 1]; ++j) {
            const u32 w = old2new[adj[j]];
            if (w >= gs) scratch.push_back(grp[w]);
          }
        }
      } else {
        for (u32 u = gs; u < start[(size_t)cv + 1]; ++u)
          for (u32 j = rp[u]; j < rp[u + 1]; ++j)
            if (adj[j] >= gs) scratch.push_back(grp[adj[j]]);
      }
      std::sort(scratch.begin(), scratch.end());
      u32 distinct = 0;
      for (size_t i = 0; i < scratch.size();) {
        size_t j = i + 1;
        while (j < scratch.size() && scratch[j] == scratch[i]) ++j;
        oa.push_back(scratch[i]);
        om.push_back((u32)(j - i));
        ++distinct;
        i = j;
      }
      up_cnt[(size_t)cv] = distinct;
    }
  }

  // stitch the per-thread upper lists into one CSR (threads own contiguous vertex ranges)
  uvec up_rp((size_t)nc + 1, 0);
  for (int64_t v = 0; v < nc; ++v) up_rp[(size_t)v + 1] = up_rp[(size_t)v] + up_cnt[(size_t)v];
  uvec up_adj(up_rp[(size_t)nc]), up_mult(up_rp[(size_t)nc]);
  for (int t = 0; t < nthreads; ++t) {
    if (buf_adj[(size_t)t].empty()) continue;
    const u32 dst = up_rp[(size_t)first_cv[(size_t)t]];
    std::memcpy(&up_adj[dst], buf_adj[(size_t)t].data(), buf_adj[(size_t)t].size() * sizeof(u32));
    std::memcpy(&up_mult[dst], buf_mult[(size_t)t].data(), buf_mult[(size_t)t].size() * sizeof(u32));
    uvec().swap(buf_adj[(size_t)t]);
    uvec().swap(buf_mult[(size_t)t]);
  }

  // transpose of the strict upper part = lower part; filling in increasing source order
  // leaves every lower list sorted ascending
  uvec low_cnt((size_t)nc + 1, 0);
  for (int64_t v = 0; v < nc; ++v)
    for (u32 e = up_rp[(size_t)v]; e < up_rp[(size_t)v + 1]; ++e)
      if (up_adj[e] != (u32)v) ++low_cnt[up_adj[e]];

  out.rp.assign((size_t)nc + 1, 0);
  for (int64_t v = 0; v < nc; ++v) {
    const u32 tot = low_cnt[(size_t)v] + up_cnt[(size_t)v];
    if (tot == 0) return HSPMV_EUNSUPPORTED; // reference reads out of bounds here (Q8)
    out.rp[(size_t)v + 1] = out.rp[(size_t)v] + tot;
  }
  out.adj.resize(out.rp[(size_t)nc]);
  out.mult.resize(out.rp[(size_t)nc]);
  uvec cursor(out.rp.begin(), out.rp.end() - 1);
  for (int64_t v = 0; v < nc; ++v)
    for (u32 e = up_rp[(size_t)v]; e < up_rp[(size_t)v + 1]; ++e) {
      const u32 t = up_adj[e];
      if (t != (u32)v) {
        out.adj[cursor[t]] = (u32)v;
        out.mult[cursor[t]] = up_mult[e];
        ++cursor[t];
      }
    }
  for (int64_t v = 0; v < nc; ++v) {
    const u32 n_up = up_cnt[(size_t)v];
    std::memcpy(&out.adj[cursor[(size_t)v]], &up_adj[up_rp[(size_t)v]], n_up * sizeof(u32));
    std::memcpy(&out.mult[cursor[(size_t)v]], &up_mult[up_rp[(size_t)v]], n_up * sizeof(u32));
  }
  return HSPMV_OK;
}

// ---- RCM on the coarse graph ---------------------------------------------------------
struct Rcm {
  const Coarse &g;
  std::vector<unsigned char> open; // 1 = not yet numbered ("mask" in the reference)
  uvec stamp;          // BFS visit epoch
  u32 epoch = 0;
  std::vector<int> order, lvl_ptr;

  explicit Rcm(const Coarse &gr) : g(gr), open((size_t)gr.n, 1), stamp((size_t)gr.n, 0) {}

  // rooted level structure over the open vertices (csrk.cpp:2426-2475)
  int levels_from(int root) {
    ++epoch;
    order.clear();
    lvl_ptr.clear();
    order.push_back(root);
    stamp[(size_t)root] = epoch;
    size_t begin = 0, end = 0;
    int levels = 0;
    for (;;) {
      begin = end;
      lvl_ptr.push_back((int)begin);
      end = order.size();
      for (size_t i = begin; i < end; ++i) {
        const int v = order[i];
        for (u32 e = g.rp[(size_t)v]; e < g.rp[(size_t)v + 1]; ++e) {
          const u32 w = g.adj[e];
          if (open[w] && stamp[w] != epoch) {
            stamp[w] = epoch;
            order.push_back((int)w);
          }
        }
      }
      ++levels;
      if (order.size() == end) break;
    }
    lvl_ptr.push_back((int)end);
    return levels;
  }

  // George-Liu pseudo-peripheral vertex (csrk.cpp:2377-2421)
  int peripheral(int root) {
    int levels = levels_from(root);
    const int cc = lvl_ptr[(size_t)levels];
    if (levels == 1 || levels == cc) return root;
    for (;;) {
      const int j1 = lvl_ptr[(size_t)levels - 1];
      int best = cc;
      root = order[(size_t)j1];
      for (int j = j1; j < cc; ++j) {
        const int v = order[(size_t)j];
        int d = 0;
        for (u32 e = g.rp[(size_t)v]; e < g.rp[(size_t)v + 1]; ++e) d += open[g.adj[e]] ? 1 : 0;
        if (d < best) {
          best = d;
          root = v;
        }
      }
      const int nl = levels_from(root);
      if (nl <= levels || levels >= cc) return root;
      levels = nl;
    }
  }
};

struct DegId { // layout of rev_deg_id_pair, csrk.h:477-490
  unsigned deg;
  unsigned id;
};
inline bool heavier_first(DegId a, DegId b) { return a.deg > b.deg; } // csrk.cpp:65-67

void rcm_order(const Coarse &g, uvec &new2old, uvec &old2new) {
  Rcm s(g);
  new2old.assign((size_t)g.n, 0);
  old2new.assign((size_t)g.n, 0);
  std::vector<int> queue;
  std::vector<DegId> cand;
  int64_t done = 0;
  for (int64_t seed = 0; seed < g.n && done < g.n; ++seed) {
    if (!s.open[(size_t)seed]) continue;
    const int root = s.peripheral((int)seed);
    queue.clear();
    queue.push_back(root);
    const int64_t first = done;
    for (size_t head = 0; head < queue.size(); ++head) {
      const int p = queue[head];
      if (!s.open[(size_t)p]) continue; // vertices are marked at dequeue, so duplicates occur
      s.open[(size_t)p] = 0;
      new2old[(size_t)done++] = (u32)p;
      cand.clear();
      for (u32 e = g.rp[(size_t)p]; e < g.rp[(size_t)p + 1]; ++e)
        if (s.open[g.adj[e]]) {
          DegId d;
          d.deg = g.mult[e];
          d.id = g.adj[e];
          cand.push_back(d);
        }
      std::sort(cand.begin(), cand.end(), heavier_first);
      for (size_t i = 0; i < cand.size(); ++i) queue.push_back((int)cand[i].id);
    }
    std::reverse(new2old.begin() + first, new2old.begin() + done);
    for (int64_t i = first; i < done; ++i) old2new[new2old[(size_t)i]] = (u32)i;
  }
}

// ---- the same RCM, level-synchronous on all host cores -----------------------------------
// At config 5 the level-1 coarse graph has 4 M vertices and ~100 M entries and the serial walk above is about
// half of the whole Band-k build.  Both of its pieces are breadth-first searches, so they can run one LEVEL at
// a time on every core and still produce the serial visit order bit for bit:
//   * rooted level structure (levels_from): a vertex of level L+1 is discovered by the frontier vertex with
//     the SMALLEST position that lists it, and within one discoverer in adjacency order.  Phase A walks the
//     frontier in contiguous chunks; every frontier position claims its undiscovered neighbours with an atomic
//     min on one word per vertex, and a chunk remembers, in walk order, the claims that succeeded.  Phase B keeps
//     the remembered claims the chunk still owns (no smaller position took them since); the chunks are
//     concatenated in chunk order.
//   * Cuthill-McKee numbering: the serial queue marks a vertex when it is DEQUEUED, so the candidate list of a
//     vertex p is "every neighbour that is dequeued after p" -- the neighbours not numbered yet plus those of
//     p's own level with a larger number -- sorted with the same std::sort call on the same sequence; the
//     vertex is numbered where it FIRST enters the queue, i.e. by its smallest-numbered neighbour, at its rank
//     in that neighbour's sorted list.  Phase A builds and sorts each p's list once and claims its unnumbered
//     entries in sorted order; phase B keeps what is still owned, as above.
// Levels below `cutoff` vertices are expanded by one thread (no fork), so a long thin graph costs what the
// serial walk costs.  tests/test_host_index.py runs both walks against the oracle and the compiled reference.
struct RcmPar {
  static constexpr u32 UNNUMBERED = 0x80000000u; // posn: numbers are < 2^31; >= 2^31 = not numbered yet (NONE, a claim, TAKEN)
  static constexpr u32 NONE = 0xFFFFFFFFu, TAKEN = 0xFFFFFFFEu;
  const Coarse &g;
  const int nt;
  const size_t cutoff;
  std::vector<unsigned char> open; // 1 = not numbered yet (the reference's mask); read by peripheral() only
  // One 32-bit word per vertex for the level structures, so that an edge costs ONE random access per phase:
  //   0            closed (numbered in an earlier component), or discovered in the current search
  //   i + 1        claimed by frontier position i in the level being expanded
  //   0xFFFFFFFF   not reached by the current search (levels_from resets the words it set when the next search starts)
  // so "claim" is a plain atomic min and leaves closed / discovered words alone.
  uvec seen;
  uvec posn; // Cuthill-McKee number, or >= UNNUMBERED: NONE / UNNUMBERED|i = claimed by number i / TAKEN
  std::vector<int> order, lvl_ptr; // order[0 .. cnt)
  size_t cnt = 0;
  // a level is cut into up to 8 chunks per thread, handed out dynamically (a core that is late does not hold the
  // level up); chunk c's finds go to cbuf[c] and the chunks are concatenated in chunk order
  std::vector<uvec> cbuf;
  std::vector<size_t> coff;

  RcmPar(const Coarse &gr, int threads, size_t cut)
      : g(gr), nt(threads), cutoff(cut), open((size_t)gr.n, 1), seen((size_t)gr.n, NONE), posn((size_t)gr.n, NONE),
        order((size_t)gr.n, 0), cbuf((size_t)threads * 8), coff((size_t)threads * 8 + 1, 0) {}

  // atomic min; true when THIS call lowered the word (the caller then holds the claim until a smaller one arrives)
  static inline bool claim_min(u32 *p, u32 v) {
    u32 cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < cur)
      if (__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return true;
    return false;
  }
  size_t chunks_of(size_t len) const { return std::max<size_t>(1, std::min<size_t>((len + 255) / 256, cbuf.size())); }
  template <typename W> static inline W peek(const W *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
  template <typename W> static inline void poke(W *p, W v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }

  // one level of the rooted level structure: frontier order[begin, end) -> appended at order[cnt ...)
  void expand_level(size_t begin, size_t end) {
    if (end - begin < cutoff || nt == 1) {
      for (size_t i = begin; i < end; ++i) {
        const int v = order[i];
        for (u32 e = g.rp[(size_t)v]; e < g.rp[(size_t)v + 1]; ++e) {
          const u32 w = g.adj[e];
          if (seen[w]) { // neither closed nor discovered in this search
            seen[w] = 0;
            order[cnt++] = (int)w;
          }
        }
      }
      return;
    }
    const size_t len = end - begin, nchunk = chunks_of(len);
    coff[0] = cnt;
#pragma omp parallel num_threads(nt)
    {
      // phase A: claims; a chunk remembers (in walk order) the neighbours it claimed and for which position
#pragma omp for schedule(dynamic, 1)
      for (int64_t c = 0; c < (int64_t)nchunk; ++c) {
        uvec &buf = cbuf[(size_t)c];
        buf.clear();
        const size_t lo = begin + len * (size_t)c / nchunk, hi = begin + len * ((size_t)c + 1) / nchunk;
        for (size_t i = lo; i < hi; ++i) {
          const int v = order[i];
          const u32 mine = (u32)i + 1;
          for (u32 e = g.rp[(size_t)v]; e < g.rp[(size_t)v + 1]; ++e) {
            const u32 w = g.adj[e];
            if (claim_min(&seen[w], mine)) {
              buf.push_back(w);
              buf.push_back(mine);
            }
          }
        }
      }
      // phase B: what a chunk claimed and still owns (no smaller position took it since) is its find, in walk order
#pragma omp for schedule(dynamic, 1)
      for (int64_t c = 0; c < (int64_t)nchunk; ++c) {
        uvec &buf = cbuf[(size_t)c];
        size_t kept = 0;
        for (size_t t = 0; t < buf.size(); t += 2)
          if (peek(&seen[buf[t]]) == buf[t + 1]) {
            poke(&seen[buf[t]], 0u);
            buf[kept++] = buf[t];
          }
        buf.resize(kept);
      }
#pragma omp single
      for (size_t c = 0; c < nchunk; ++c) coff[c + 1] = coff[c] + cbuf[c].size();
#pragma omp for schedule(dynamic, 1) nowait
      for (int64_t c = 0; c < (int64_t)nchunk; ++c) {
        const uvec &buf = cbuf[(size_t)c];
        for (size_t t = 0; t < buf.size(); ++t) order[coff[(size_t)c] + t] = (int)buf[t];
      }
    }
    cnt = coff[nchunk];
  }

  // un-discover the vertices of the previous search (they are all still open)
  void forget() {
    const int64_t c = (int64_t)cnt;
#pragma omp parallel for schedule(static) num_threads(nt) if (c > 65536)
    for (int64_t i = 0; i < c; ++i) seen[(size_t)order[(size_t)i]] = NONE;
    cnt = 0;
  }

  int levels_from(int root) {
    lvl_ptr.clear();
    forget();
    order[cnt++] = root;
    seen[(size_t)root] = 0;
    size_t begin = 0, end = 0;
    int levels = 0;
    for (;;) {
      begin = end;
      lvl_ptr.push_back((int)begin);
      end = cnt;
      expand_level(begin, end);
      ++levels;
      if (cnt == end) break;
    }
    lvl_ptr.push_back((int)end);
    return levels;
  }

  int peripheral(int root) { // identical to Rcm::peripheral, on the parallel level structure
    int levels = levels_from(root);
    const int cc = lvl_ptr[(size_t)levels];
    if (levels == 1 || levels == cc) return root;
    for (;;) {
      const int j1 = lvl_ptr[(size_t)levels - 1];
      int best = cc;
      root = order[(size_t)j1];
      for (int j = j1; j < cc; ++j) {
        const int v = order[(size_t)j];
        int d = 0;
        for (u32 e = g.rp[(size_t)v]; e < g.rp[(size_t)v + 1]; ++e) d += open[g.adj[e]] ? 1 : 0;
        if (d < best) {
          best = d;
          root = v;
        }
      }
      const int nl = levels_from(root);
      if (nl <= levels || levels >= cc) return root;
      levels = nl;
    }
  }

  // candidates of the vertex numbered i = the neighbours dequeued after it (unnumbered ones have posn >= 2^31 > i;
  // vertices of earlier components have smaller numbers), sorted as the serial walk sorts them
  inline void candidates(u32 p, u32 i, std::vector<DegId> &cand) const {
    cand.clear();
    for (u32 e = g.rp[p]; e < g.rp[(size_t)p + 1]; ++e) {
      const u32 w = g.adj[e];
      if (peek(&posn[w]) > i) {
        DegId d;
        d.deg = g.mult[e];
        d.id = w;
        cand.push_back(d);
      }
    }
    std::sort(cand.begin(), cand.end(), heavier_first);
  }

  // Cuthill-McKee numbers first, first+1, ... of the component of `root` into q[]; returns the next free number
  int64_t number_component(int root, int64_t first, uvec &q) {
    forget();
    size_t begin = (size_t)first, end = begin + 1;
    q[begin] = (u32)root;
    posn[(size_t)root] = (u32)begin;
    std::vector<DegId> cand0;
    while (begin < end) {
      size_t next = end;
      if (end - begin < cutoff || nt == 1) {
        for (size_t i = begin; i < end; ++i) {
          candidates(q[i], (u32)i, cand0);
          for (size_t c = 0; c < cand0.size(); ++c)
            if (posn[cand0[c].id] == NONE) {
              posn[cand0[c].id] = (u32)next;
              q[next++] = cand0[c].id;
            }
        }
      } else {
        const size_t len = end - begin, nchunk = chunks_of(len);
        coff[0] = end;
#pragma omp parallel num_threads(nt)
        {
          // phase A: every p builds and sorts its list once and claims its unnumbered entries in sorted order; the chunk
          // remembers what it claimed (numbers are smaller than any claim, so an atomic min leaves numbered vertices alone)
          std::vector<DegId> cand;
#pragma omp for schedule(dynamic, 1)
          for (int64_t c = 0; c < (int64_t)nchunk; ++c) {
            uvec &buf = cbuf[(size_t)c];
            buf.clear();
            const size_t lo = begin + len * (size_t)c / nchunk, hi = begin + len * ((size_t)c + 1) / nchunk;
            for (size_t i = lo; i < hi; ++i) {
              candidates(q[i], (u32)i, cand);
              const u32 mine = UNNUMBERED | (u32)i;
              for (size_t k = 0; k < cand.size(); ++k)
                if (claim_min(&posn[cand[k].id], mine)) {
                  buf.push_back(cand[k].id);
                  buf.push_back(mine);
                }
            }
          }
          // phase B: keep what is still owned, in (p, rank in p's sorted list) order
#pragma omp for schedule(dynamic, 1)
          for (int64_t c = 0; c < (int64_t)nchunk; ++c) {
            uvec &buf = cbuf[(size_t)c];
            size_t kept = 0;
            for (size_t t = 0; t < buf.size(); t += 2)
              if (peek(&posn[buf[t]]) == buf[t + 1]) {
                poke(&posn[buf[t]], TAKEN);
                buf[kept++] = buf[t];
              }
            buf.resize(kept);
          }
#pragma omp single
          for (size_t c = 0; c < nchunk; ++c) coff[c + 1] = coff[c] + cbuf[c].size();
#pragma omp for schedule(dynamic, 1) nowait
          for (int64_t c = 0; c < (int64_t)nchunk; ++c) {
            const uvec &buf = cbuf[(size_t)c];
            for (size_t t = 0; t < buf.size(); ++t) {
              q[coff[(size_t)c] + t] = buf[t];
              posn[buf[t]] = (u32)(coff[(size_t)c] + t);
            }
          }
        }
        next = coff[nchunk];
      }
      begin = end;
      end = next;
    }
    // the component is numbered: close it for the searches of the following components
    const int64_t done = (int64_t)end;
#pragma omp parallel for schedule(static) num_threads(nt) if (done - first > 65536)
    for (int64_t i = first; i < done; ++i) {
      open[q[(size_t)i]] = 0;
      seen[q[(size_t)i]] = 0;
    }
    return done;
  }
};

void rcm_order_parallel(const Coarse &g, int nt, size_t cutoff, uvec &new2old, uvec &old2new) {
  RcmPar s(g, nt, cutoff);
  new2old.assign((size_t)g.n, 0);
  old2new.assign((size_t)g.n, 0);
  int64_t done = 0;
  for (int64_t seed = 0; seed < g.n && done < g.n; ++seed) {
    if (!s.open[(size_t)seed]) continue;
    const int root = s.peripheral((int)seed);
    const int64_t first = done;
    done = s.number_component(root, first, new2old);
    std::reverse(new2old.begin() + first, new2old.begin() + done);
#pragma omp parallel for schedule(static) num_threads(nt) if (done - first > 65536)
    for (int64_t i = first; i < done; ++i) old2new[new2old[(size_t)i]] = (u32)i;
  }
}

// serial walk for small graphs, level-synchronous walk on all cores for large ones (HSPMV_RCM=serial|parallel and
// HSPMV_RCM_CUTOFF=<frontier size> override the choice; the tests use them to run both on the same graphs)
void rcm_order_auto(const Coarse &g, uvec &new2old, uvec &old2new) {
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  const char *mode = std::getenv("HSPMV_RCM");
  const char *cut = std::getenv("HSPMV_RCM_CUTOFF");
  size_t cutoff = cut ? (size_t)std::strtoull(cut, nullptr, 10) : 2048;
  if (cutoff < 1) cutoff = 1;
  bool par = nt > 1 && g.n >= (1 << 16);
  if (g.n >= (int64_t)RcmPar::UNNUMBERED) par = false, mode = nullptr; // numbers and claims share one 32-bit word
  if (mode && std::strcmp(mode, "serial") == 0) par = false;
  if (mode && std::strcmp(mode, "parallel") == 0) par = true;
  if (par)
    rcm_order_parallel(g, nt, cutoff, new2old, old2new);
  else
    rcm_order(g, new2old, old2new);
}

// relabel + sort adjacency by new id (csrk.cpp:3012-3115)
void relabel(Coarse &g, const uvec &new2old, const uvec &old2new) {
  Coarse src;
  src.rp.swap(g.rp);
  src.adj.swap(g.adj);
  src.mult.swap(g.mult);
  g.rp.assign((size_t)g.n + 1, 0);
  for (int64_t i = 0; i < g.n; ++i) {
    const u32 o = new2old[(size_t)i];
    g.rp[(size_t)i + 1] = g.rp[(size_t)i] + (src.rp[o + 1] - src.rp[o]);
  }
  g.adj.resize(src.adj.size());
  g.mult.resize(src.mult.size());
#pragma omp parallel
  {
    std::vector<std::pair<u32, u32>> row;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < g.n; ++i) {
      const u32 o = new2old[(size_t)i];
      row.clear();
      for (u32 e = src.rp[o]; e < src.rp[o + 1]; ++e)
        row.push_back(std::make_pair(old2new[src.adj[e]], src.mult[e]));
      std::sort(row.begin(), row.end()); // ids are distinct after dedup
      for (size_t t = 0; t < row.size(); ++t) {
        g.adj[g.rp[(size_t)i] + t] = row[t].first;
        g.mult[g.rp[(size_t)i] + t] = row[t].second;
      }
    }
  }
}

// csrk.cpp:1148-1224: group starts rewritten in permuted order; finer permutation composed.  An EMPTY perm_fine
// stands for the identity (level 0, csrk.cpp:886-890: the rows themselves), which is then never materialised.
void expand_level(uvec &map, const uvec &perm_coarse, uvec &perm_fine) {
  const size_t nc = perm_coarse.size();
  const bool identity = perm_fine.empty();
  uvec new_map(nc + 1, 0), composed(identity ? (size_t)map[nc] : perm_fine.size());
  for (size_t j = 0; j < nc; ++j) {
    const u32 o = perm_coarse[j];
    new_map[j + 1] = new_map[j] + (map[o + 1] - map[o]);
  }
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < (int64_t)nc; ++j) {
    const u32 o = perm_coarse[(size_t)j];
    u32 dst = new_map[(size_t)j];
    if (identity)
      for (u32 t = map[o]; t < map[o + 1]; ++t) composed[dst++] = t;
    else
      for (u32 t = map[o]; t < map[o + 1]; ++t) composed[dst++] = perm_fine[t];
  }
  for (size_t t = new_map[nc]; t < composed.size(); ++t) composed[t] = 0; // rows no group covers (trailing empty rows)
  map.swap(new_map);
  perm_fine.swap(composed);
}

} // namespace

// csrk.cpp:896-897: unsigned * long / long, truncated to int, compared as unsigned
void first_touch(void *p, size_t bytes) {
  volatile unsigned char *b = (volatile unsigned char *)p;
  const int64_t pages = (int64_t)((bytes + 4095) / 4096);
#pragma omp parallel for schedule(static) if (pages > 1024)
  for (int64_t i = 0; i < pages; ++i) b[(size_t)i * 4096] = 0;
}

u32 bandk_threshold(int size, int64_t nnz, int64_t n) {
  return (u32)(int)((long)(unsigned)size * (long)nnz / (long)n);
}

void bandk_group_rows(const u32 *rp, int64_t n, u32 T, uvec &start) { group_starts(rp, n, T, start); }

// for the tests: both statements of the grouping rule side by side
void bandk_group_rows_literal(const u32 *rp, int64_t n, u32 T, uvec &start, uvec &grp) {
  group_rows_literal(rp, n, T, start, grp);
}
void bandk_group_rows_fast(const u32 *rp, int64_t n, u32 T, uvec &start, uvec &grp) {
  group_rows(rp, n, T, start, grp);
}

// Levels 1 .. k-1.  `first` (optional) is the level-1 coarse graph with its group starts in
// `first_map`, built elsewhere (bandk_gpu.cu); without it level 1 is built here from (rp, adj).
static int bandk_levels(int64_t n, int64_t nnz, const u32 *rp, const u32 *adj, int k, const int *sizes,
                        CoarseGraph *first, uvec *first_map, uvec &perm,
                        uvec &sr_ptr, uvec &ssr_ptr) {
  std::vector<uvec> perms((size_t)k), maps((size_t)k);
  Coarse level[3];
  uvec fused_n2o, fused_o2n, fused_rp; // level l in its old numbering + the renumbering, while level l+1 is coarsened

  const u32 *cur_rp = rp, *cur_adj = adj;
  int64_t cur_n = n, cur_nnz = nnz;
  for (int lvl = 1; lvl < k; ++lvl) {
    if (lvl == 1 && first) {
      level[1].n = first->n;
      level[1].rp.swap(first->rp);
      level[1].adj.swap(first->adj);
      level[1].mult.swap(first->mult);
      maps[1].swap(*first_map);
    } else {
      const u32 T = bandk_threshold(sizes[lvl - 1], cur_nnz, cur_n);
      uvec grp;
      {
        Phase ph("group rows");
        group_rows(cur_rp, cur_n, T, maps[(size_t)lvl], grp);
      }
      Phase ph("coarse graph (host)");
      int rc;
      if (lvl > 1 && !fused_n2o.empty()) {
        rc = build_coarse(level[lvl - 1].rp.data(), level[lvl - 1].adj.data(), maps[(size_t)lvl], grp, level[lvl], fused_n2o.data(),
                          fused_o2n.data());
        perms[(size_t)lvl - 1].swap(fused_n2o); // back where the uncoarsening expects it
      } else {
        rc = build_coarse(cur_rp, cur_adj, maps[(size_t)lvl], grp, level[lvl]);
      }
      if (rc) return rc;
    }
    uvec o2n;
    {
      Phase ph("RCM of the coarse graph");
      rcm_order_auto(level[lvl], perms[(size_t)lvl], o2n);
    }
    if (lvl + 1 < k) { // the next level coarsens the renumbered graph
      cur_n = level[lvl].n;
      cur_nnz = (int64_t)level[lvl].adj.size();
      const char *lit = std::getenv("HSPMV_BANDK_RELABEL"); // "1": relabel first, then coarsen (the literal two steps; tests)
      if (lit && lit[0] == '1') {
        Phase ph("relabel");
        relabel(level[lvl], perms[(size_t)lvl], o2n);
        cur_rp = level[lvl].rp.data();
        cur_adj = level[lvl].adj.data();
        fused_n2o.clear();
      } else { // renumbering applied on the fly by build_coarse: only the row pointer of the new order is built
        Phase ph("row pointer in the new order");
        fused_n2o.swap(perms[(size_t)lvl]);
        fused_o2n.swap(o2n);
        fused_rp.resize((size_t)cur_n + 1);
        fused_rp[0] = 0;
        for (int64_t i = 0; i < cur_n; ++i) {
          const u32 o = fused_n2o[(size_t)i];
          fused_rp[(size_t)i + 1] = fused_rp[(size_t)i] + (level[lvl].rp[o + 1] - level[lvl].rp[o]);
        }
        cur_rp = fused_rp.data();
        cur_adj = nullptr; // (not addressable in the new order; build_coarse goes through new2old / old2new)
      }
    }
  }

  // level-(k-1) ... level-1: rewrite the maps, compose down to rows (csrk.cpp:1015-1034)
  Phase ph("uncoarsen");
  if (k == 3) expand_level(maps[2], perms[2], perms[1]);
  perms[0].clear(); // level 0 is the identity: not materialised ...
  if ((int64_t)maps[1].back() != n) { // ... unless trailing empty rows are left uncovered (the callers report that case)
    perms[0].resize((size_t)n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) perms[0][(size_t)i] = (u32)i;
  }
  expand_level(maps[1], perms[1], perms[0]);

  perm.swap(perms[0]);
  sr_ptr.swap(maps[1]);
  if (k == 3)
    ssr_ptr.swap(maps[2]);
  else
    ssr_ptr.clear();
  return HSPMV_OK;
}

int bandk_build(int64_t n, int64_t nnz, const u32 *rp, const u32 *adj, int k, const int *sizes,
                uvec &perm, uvec &sr_ptr, uvec &ssr_ptr) {
  if (k < 2 || k > 3 || n <= 0 || !rp || !adj || !sizes) return HSPMV_EINVAL;
  if (rp[n] != (u32)nnz) return HSPMV_EINVAL;
  return bandk_levels(n, nnz, rp, adj, k, sizes, nullptr, nullptr, perm, sr_ptr, ssr_ptr);
}

int bandk_finish(int64_t n, int k, const int *sizes, CoarseGraph &level1, uvec &map1,
                 uvec &perm, uvec &sr_ptr, uvec &ssr_ptr) {
  if (k < 2 || k > 3 || n <= 0 || !sizes) return HSPMV_EINVAL;
  return bandk_levels(n, 0, nullptr, nullptr, k, sizes, &level1, &map1, perm, sr_ptr, ssr_ptr);
}

template <typename T>
static void reorder_impl(int64_t n, const u32 *perm, const u32 *rp, const u32 *ci, const T *v,
                         u32 *rp2, u32 *ci2, T *v2) {
  uvec fwd((size_t)n);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) fwd[perm[i]] = (u32)i;
  rp2[0] = 0;
  for (int64_t i = 0; i < n; ++i) rp2[i + 1] = rp2[i] + (rp[perm[i] + 1] - rp[perm[i]]);
#pragma omp parallel
  {
    std::vector<std::pair<u32, u32>> row; // (new column, old position)
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; ++i) {
      const u32 o = perm[i];
      const u32 len = rp[o + 1] - rp[o];
      row.resize(len);
      bool sorted = true;
      for (u32 t = 0; t < len; ++t) {
        row[t].first = fwd[ci[rp[o] + t]];
        row[t].second = rp[o] + t;
        if (t && row[t].first < row[t - 1].first) sorted = false;
      }
      if (!sorted) // columns of a row are unique, so any sort gives the reference's order
        std::sort(row.begin(), row.end());
      const u32 dst = rp2[i];
      for (u32 t = 0; t < len; ++t) {
        ci2[dst + t] = row[t].first;
        v2[dst + t] = v[row[t].second];
      }
    }
  }
}

void reorder_f32(int64_t n, const u32 *perm, const u32 *rp, const u32 *ci, const float *v,
                 u32 *rp2, u32 *ci2, float *v2) {
  reorder_impl<float>(n, perm, rp, ci, v, rp2, ci2, v2);
}
void reorder_f64(int64_t n, const u32 *perm, const u32 *rp, const u32 *ci, const double *v,
                 u32 *rp2, u32 *ci2, double *v2) {
  reorder_impl<double>(n, perm, rp, ci, v, rp2, ci2, v2);
}

} // namespace hspmv
