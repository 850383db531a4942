// Sparse triangular solve L x = b on the level-set-ordered matrix (replaces CSRk_Graph::lowerSTS,
// k == 2, spmv-csrk/csrk.cpp:79-89, with incomplete_choloskey :708-789 and compute_b :791-809).
//
// L is not materialised: after hspmv_build_sts the handle holds A' = P A P^T with sorted rows, and diag[r] is the
// position of row r's diagonal entry, so L's row r is the entries [rp[r], diag[r]] of A' -- the diagonal last, as in
// the reference's val_L / adj_L.  Rows of one pack (one level) do not depend on each other; packs run in order:
//   * the leading packs are small (levels are sorted by size): CTA 0 (1024 threads) runs them alone with
//     __syncthreads between them and plain loads of x (its own SM wrote them); the other CTAs wait at one grid barrier;
//   * the remaining packs are spread over the whole grid (one CTA per SM: a grid barrier costs less with fewer CTAs),
//     VEC lanes per row, one cooperative grid barrier per pack; x entries written in earlier packs are read with
//     ld.global.cg (L2): another SM wrote them during this launch.
#pragma once
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace hspmv {

// diag[r] = position of the entry (r, r); *bad = 1 when a row has none ("Self edge not found", csrk.cpp:731-733)
static __global__ void sts_diag_kernel(int64_t n, const int *__restrict__ rp, const int *__restrict__ ci, int *__restrict__ diag,
                                       int *__restrict__ bad) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    int lo = rp[r], hi = rp[r + 1];
    while (lo < hi) { // first position with column >= r
      const int mid = (lo + hi) >> 1;
      if (ci[mid] < (int)r) lo = mid + 1; else hi = mid;
    }
    if (lo < rp[r + 1] && ci[lo] == (int)r) diag[r] = lo;
    else {
      diag[r] = rp[r];
      *bad = 1;
    }
  }
}

// b = L * 1, one accumulator per row, left to right (compute_b): bit-identical to the reference's loop
template <typename T>
__global__ void sts_b_kernel(int64_t n, const int *__restrict__ rp, const int *__restrict__ diag, const T *__restrict__ val,
                             T *__restrict__ b) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    T t = 0;
    for (int j = rp[r]; j <= diag[r]; ++j) t += val[j] * (T)1.0;
    b[r] = t;
  }
}

template <typename T, int VEC, bool LOCAL>
__device__ __forceinline__ void sts_rows(int row_lo, int row_hi, int group, int n_groups, int sub, const int *__restrict__ rp,
                                         const int *__restrict__ diag, const int *__restrict__ ci, const T *__restrict__ val,
                                         const T *__restrict__ b, T *x) {
  // warp-uniform trip count (full-mask shuffles): groups of one warp advance together
  const int per_warp = 32 / VEC;
  for (int base = row_lo + (group / per_warp) * per_warp; base < row_hi; base += n_groups) {
    const int r = base + group % per_warp;
    const bool active = r < row_hi;
    T acc = 0, rhs = 0, dia = 1;
    int rs = 0, d = 0;
    if (active) {
      rs = rp[r];
      d = diag[r];
      rhs = b[r];   // both independent of x: in flight while the dependencies are gathered
      dia = val[d];
    }
    // four entries per lane per round, all column / value loads of a round issued before its x gathers: a row costs
    // ceil(lower entries / (4 * VEC)) rounds of latency, not one per entry (warp-uniform trip count)
    for (int j0 = rs + sub; __any_sync(0xffffffffu, j0 < d); j0 += 4 * VEC) {
      int c[4];
      T v[4], xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = j0 + u * VEC < d;
        c[u] = ok ? ci[j0 + u * VEC] : 0;
        v[u] = ok ? val[j0 + u * VEC] : (T)0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        // LOCAL: every x entry read here was written by THIS CTA (its own SM's L1 is coherent with its own stores);
        // otherwise another SM wrote it during this launch: read it from L2
        const bool ok = j0 + u * VEC < d;
        xv[u] = ok ? (LOCAL ? x[c[u]] : __ldcg(x + c[u])) : (T)0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += v[u] * xv[u];
    }
    acc = group_sum<T, VEC>(acc);
    if (active && sub == 0) x[r] = (rhs - acc) / dia;
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(1024) sts_solve_kernel(int n_packs, int solo_packs, const int *__restrict__ pack_ptr,
                                                        const int *__restrict__ rp, const int *__restrict__ diag,
                                                        const int *__restrict__ ci, const T *__restrict__ val,
                                                        const T *__restrict__ b, T *x) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  const int sub = threadIdx.x & (VEC - 1);
  if (blockIdx.x == 0) { // the small leading packs: one CTA, block barriers
    const int group = threadIdx.x / VEC, n_groups = blockDim.x / VEC;
    for (int p = 0; p < solo_packs; ++p) {
      sts_rows<T, VEC, true>(pack_ptr[p], pack_ptr[p + 1], group, n_groups, sub, rp, diag, ci, val, b, x);
      __syncthreads();
    }
  }
  if (blockIdx.x == 0) __threadfence();
  if (solo_packs > 0 && solo_packs < n_packs) grid.sync();
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / VEC, n_groups = gridDim.x * blockDim.x / VEC;
  for (int p = solo_packs; p < n_packs; ++p) {
    sts_rows<T, VEC, false>(pack_ptr[p], pack_ptr[p + 1], group, n_groups, sub, rp, diag, ci, val, b, x);
    if (p + 1 < n_packs) grid.sync();
  }
}

// ---- the solve without barriers ------------------------------------------------------------------------------
// Rows are dealt to groups of VEC lanes in increasing order (group g takes rows g, g + G, g + 2G, ...; every group
// of the launch is resident: cooperative launch).  A row waits for exactly the x entries it reads: ready[c] holds the
// epoch of the solve that last wrote x[c]; a lane spins (bounded) on ready[c] with an acquire load, reads x[c] from L2,
// and after the group's sum one lane writes x[r], fences, and publishes ready[r] = epoch.  No deadlock: L is lower
// triangular in this numbering, so the smallest unfinished row never waits, and the group that owns it is AT it
// (its earlier rows are smaller).  Consecutive packs overlap: a row starts as soon as ITS dependencies are done, not
// when the whole previous level is -- the level-scheduled kernel above pays a block or grid barrier (~2-3.5 us) per
// pack, this one a flag round trip through L2 per dependency chain link.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) sts_flags_kernel(int n, int epoch, const int *__restrict__ rp, const int *__restrict__ diag,
                                                        const int *__restrict__ ci, const T *__restrict__ val,
                                                        const T *__restrict__ b, T *x, int *ready, int *timed_out) {
  const int sub = threadIdx.x & (VEC - 1);
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / VEC, n_groups = gridDim.x * blockDim.x / VEC;
  const unsigned gmask = VEC == 32 ? 0xffffffffu : (((1u << VEC) - 1u) << ((threadIdx.x & 31) & ~(VEC - 1)));
  for (int r = group; r < n; r += n_groups) {
    const int rs = rp[r], d = diag[r];
    const T rhs = b[r], dia = val[d];
    T acc = 0;
    for (int j0 = rs + sub; j0 < d; j0 += 4 * VEC) { // (per-group trip count: the shuffles below name the group only)
      int c[4];
      T v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = j0 + u * VEC < d;
        c[u] = ok ? ci[j0 + u * VEC] : -1;
        v[u] = ok ? val[j0 + u * VEC] : (T)0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (c[u] < 0) continue;
        int seen;
        const long long t0 = clock64();
        for (;;) {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(ready + c[u]) : "memory");
          if (seen == epoch) break;
          if (clock64() - t0 > 2000000000LL) { // ~1 s: never in a correct run; do not hang the device
            atomicExch(timed_out, 1);
            break;
          }
          __nanosleep(32); // a spinning lane must not crowd out the groups whose rows it is waiting for
        }
        acc += v[u] * __ldcg(x + c[u]);
      }
    }
#pragma unroll
    for (int o = VEC >> 1; o > 0; o >>= 1) acc += __shfl_down_sync(gmask, acc, o, VEC);
    if (sub == 0) {
      x[r] = (rhs - acc) / dia;
      __threadfence();
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(ready + r), "r"(epoch) : "memory");
    }
  }
}

} // namespace hspmv
