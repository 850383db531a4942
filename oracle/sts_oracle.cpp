// ============================================================================
// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's sparse triangular solve on
// CSR-k with k = 2 ("STS-2"), level-set (LS) ordering.  SURVEY.md section 8, row f4.
//
// Nothing in the product links, loads or calls this file.
//
// Parity status: PINNED against the reference itself (oracle/_ref/libref_csrk_*.so: the
// unmodified /root/reference/spmv-csrk/csrk.cpp driven through oracle/ref_shim_csrk.cpp's
// ref_sts_* entry points) in tests/test_sts.py.  Scope: kernel "STS", order "LS", k = 2 --
// the path of BAND_k::preprocessingForSTS that needs neither Boost (COLOR ordering,
// csrk.cpp:2946-3009, compile-only stub here) nor the matching coarseners; k >= 3 (STS on a
// coarsened graph, csrk.cpp:1693-2020) is not restated.
//
// What is restated (file:line relative to /root/reference/spmv-csrk/csrk.cpp):
//   sts_oracle_levels   :2704-2816  find_levels (literally, with the deletion marks), then
//                       :1612-1666  levels sorted by size with std::sort + compare_deg_id_pair
//                                   (:57-59, non-stable: the tie order is libstdc++'s, so the same
//                                   call on the same element layout is made here), permBigG =
//                                   the concatenated levels, packsPointer = their bounds
//   sts_oracle_lower    :708-789    incomplete_choloskey: L = entries with column <= row of the
//                                   REORDERED matrix (the diagonal is the last entry of a row)
//   sts_oracle_b_*      :791-809    compute_b: b = L * 1, one accumulator per row, left to right
//   sts_oracle_solve_*  :79-89      lowerSTS, k == 2: packs in order, rows of a pack in parallel
//
// Built with the SAME floating-point flags as the reference library in oracle/_ref (-O3 -march=x86-64-v3, GNU's
// default -ffp-contract=fast: `t += v * x` is one fused multiply-add), unlike spmv_oracle.c: the solve divides, so its
// x is not exactly representable and a fused and an unfused accumulation differ in the last bits.  With equal flags
// the restated loop and the reference's loop give bit-identical x, which is what tests/test_sts.py asserts.
// ============================================================================
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
const int VERTEX_MARKER = -10; // csrk.h:55
struct degree_id_pair {        // csrk.h:84-98
  int deg;
  int id;
  degree_id_pair() : deg(0), id(0) {}
  degree_id_pair(int a, int b) : deg(a), id(b) {}
};
bool compare_deg_id_pair(degree_id_pair p1, degree_id_pair p2) { return p1.deg < p2.deg; }
} // namespace

extern "C" {

// perm[n] (new -> old), pack_ptr[n+1 slots] (n_packs + 1 used).  Returns 0, or 1 when the level
// search stalls ("Number of vertices did not change in level", csrk.cpp:2797-2800: a row without a
// stored diagonal, or a structurally unsymmetric matrix).
int sts_oracle_levels(long n, long nnz, const int *num_edges, const int *adj, unsigned *perm, int *pack_ptr, int *n_packs) {
  std::vector<int> adj_copy(adj, adj + nnz), levelSetPtr((size_t)n + 2), candLevelSet((size_t)n), adjLevelSet((size_t)n);
  std::vector<char> processedVerts((size_t)n, 0), notCandidateVerts((size_t)n, 0);
  int indexVer = 0, candIndex = 0, numLevel = 0;
  levelSetPtr[0] = 0;
  bool stalled = false;
  while (indexVer < n) {
    numLevel++;
    if (numLevel == 1) {
      for (int iRow = 0; iRow < n; iRow++)
        if (num_edges[iRow] < num_edges[iRow + 1] && adj_copy[num_edges[iRow]] == iRow) {
          adjLevelSet[indexVer] = iRow;
          processedVerts[iRow] = 1;
          notCandidateVerts[iRow] = 1;
          indexVer++;
        }
    } else {
      for (int iRow = 0; iRow < candIndex; iRow++) {
        const int v = candLevelSet[iRow];
        bool firstPositiveNeighbor = true;
        for (int jInd = num_edges[v]; jInd < num_edges[v + 1]; jInd++) {
          if (adj_copy[jInd] >= 0) {
            if (adj_copy[jInd] == v) {
              adjLevelSet[indexVer] = v;
              indexVer++;
              processedVerts[v] = 1;
              firstPositiveNeighbor = false;
            } else {
              notCandidateVerts[v] = 0;
              firstPositiveNeighbor = false;
            }
          }
          if (!firstPositiveNeighbor) break;
        }
      }
    }
    levelSetPtr[numLevel] = indexVer;
    candIndex = 0;
    for (int i = levelSetPtr[numLevel - 1]; i < levelSetPtr[numLevel]; i++) {
      const int g = adjLevelSet[i];
      for (int jInd = num_edges[g]; jInd < num_edges[g + 1]; jInd++) {
        const int nb = adj[jInd];
        if (!processedVerts[nb]) {
          if (!notCandidateVerts[nb]) {
            candLevelSet[candIndex] = nb;
            notCandidateVerts[nb] = 1;
            candIndex++;
          }
          for (int j2 = num_edges[nb]; j2 < num_edges[nb + 1]; j2++)
            if (adj_copy[j2] == g) adj_copy[j2] = VERTEX_MARKER;
        }
      }
    }
    if (levelSetPtr[numLevel] - levelSetPtr[numLevel - 1] == 0) {
      stalled = true;
      break;
    }
  }
  if (stalled || indexVer != n) return 1;
  // csrk.cpp:1617-1654: levels in increasing order of size
  std::vector<degree_id_pair> levelVecSize((size_t)numLevel);
  for (int i = 0; i < numLevel; i++) levelVecSize[i] = degree_id_pair(levelSetPtr[i + 1] - levelSetPtr[i], i);
  std::sort(levelVecSize.begin(), levelVecSize.begin() + numLevel, compare_deg_id_pair);
  int cumulativeIndex = 0;
  pack_ptr[0] = 0;
  for (int i = 0; i < numLevel; i++) {
    const int old = levelVecSize[i].id;
    for (int j = levelSetPtr[old]; j < levelSetPtr[old + 1]; j++) perm[cumulativeIndex++] = (unsigned)adjLevelSet[j];
    pack_ptr[i + 1] = cumulativeIndex;
  }
  *n_packs = numLevel;
  return 0;
}

// L of an (already reordered) matrix: rows keep their entries with column <= row.  l_rp[n+1]; l_ci / l_pos sized nnz
// (l_pos[e] = position of L's entry e in the matrix arrays, so that the caller copies the values of either type).
// Returns the number of entries of L, or -1 when a row has no diagonal ("Self edge not found", csrk.cpp:731-733).
long sts_oracle_lower(long n, const unsigned *r, const unsigned *c, unsigned *l_rp, unsigned *l_ci, long *l_pos) {
  long cnt = 0;
  for (long i = 0; i < n; ++i) {
    l_rp[i] = (unsigned)cnt;
    bool self = false;
    for (unsigned j = r[i]; j < r[i + 1]; ++j) {
      if (c[j] < (unsigned)i || c[j] == (unsigned)i) {
        l_ci[cnt] = c[j];
        l_pos[cnt] = (long)j;
        ++cnt;
        self = self || c[j] == (unsigned)i;
      }
    }
    if (!self) return -1;
  }
  l_rp[n] = (unsigned)cnt;
  return cnt;
}

#define DEFINE_STS(T, SFX)                                                                                        \
  void sts_oracle_b_##SFX(long n, const unsigned *l_rp, const unsigned *l_ci, const T *l_val, T *b) {              \
    (void)l_ci;                                                                                                   \
    for (long i = 0; i < n; ++i) {                                                                                \
      T t = 0;                                                                                                    \
      for (unsigned j = l_rp[i]; j < l_rp[i + 1]; ++j) t += l_val[j] * (T)1.0;                                    \
      b[i] = t;                                                                                                   \
    }                                                                                                             \
  }                                                                                                               \
  void sts_oracle_solve_##SFX(long n, int n_packs, const int *pack_ptr, const unsigned *l_rp, const unsigned *l_ci, \
                              const T *l_val, const T *b, T *x) {                                                 \
    (void)n;                                                                                                      \
    for (int p = 0; p < n_packs; ++p) {                                                                           \
      _Pragma("omp parallel for schedule(runtime)")                                                               \
      for (int j = pack_ptr[p]; j < pack_ptr[p + 1]; ++j) {                                                       \
        T t = 0;                                                                                                  \
        for (unsigned k = l_rp[j]; k < l_rp[j + 1] - 1; ++k) t += l_val[k] * x[l_ci[k]];                          \
        x[j] = (b[j] - t) / l_val[l_rp[j + 1] - 1];                                                               \
      }                                                                                                           \
    }                                                                                                             \
  }
DEFINE_STS(float, f32)
DEFINE_STS(double, f64)

} // extern "C"
