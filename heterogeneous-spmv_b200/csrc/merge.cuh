// CSR_MERGE: nnz-balanced SpMV for power-law matrices (BASELINE.json config 4: R-MAT).
//
// The nonzeros are cut into tiles of `tile` consecutive entries regardless of row boundaries
// (perfect balance, perfectly coalesced 128-bit streaming of col/val).  tile_row[t] -- the row
// that owns nonzero t*tile -- is precomputed once per matrix (merge-path coordinates;
// partition.cpp:tile_rows is the host restatement the tests diff against).
// Every CTA owns a CONTIGUOUS range of tiles, so a row that straddles tiles is carried from tile
// to tile in shared memory; only the (at most one) row that straddles a CTA's last tile is
// written to carry_row/carry_val and added by merge_fixup_kernel in CTA order -> deterministic.
//
// Per tile:  phase 1  prod[i] = val[i] * x[col[i]]        (all threads, 4 entries per step)
//            phase 2  row sums over prod[] by groups of `vec` lanes (vec chosen per tile from
//                     its mean row length), first/last row combined with the carries.
#pragma once
#include "kernels.cuh"

namespace hspmv {

__global__ void tile_rows_kernel(int64_t m, const int *__restrict__ rp, int64_t tile,
                                 int64_t n_tiles, int *__restrict__ tile_row) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t <= n_tiles;
       t += (int64_t)gridDim.x * blockDim.x) {
    if (t == n_tiles) {
      tile_row[t] = (int)m;
      continue;
    }
    const int64_t e = t * tile;
    // last row r with rp[r] <= e
    int64_t lo = 0, hi = m + 1; // first index with rp[idx] > e lies in (lo, hi]
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)rp[mid] <= e) lo = mid + 1; else hi = mid;
    }
    tile_row[t] = (int)(lo - 1);
  }
}

template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS)
csr_merge_kernel(int64_t m, int64_t nnz, int tile, int64_t n_tiles,
                 const int *__restrict__ tile_row, const int *__restrict__ rp,
                 const int *__restrict__ ci, const T *__restrict__ val, const T *__restrict__ x,
                 T *__restrict__ y, int *__restrict__ carry_row, T *__restrict__ carry_val) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  T *prod = reinterpret_cast<T *>(merge_smem);
  __shared__ T carry[2];
  __shared__ int carry_pending[2];

  const int64_t t0 = n_tiles * blockIdx.x / gridDim.x;
  const int64_t t1 = n_tiles * (blockIdx.x + 1) / gridDim.x;
  if (threadIdx.x < 2) {
    carry[threadIdx.x] = 0;
    carry_pending[threadIdx.x] = -1;
  }
  __syncthreads();

  for (int64_t t = t0; t < t1; ++t) {
    const int64_t lo = t * tile;
    const int64_t hi = (lo + tile < nnz) ? lo + tile : nnz;
    const int par = (int)((t - t0) & 1);

    // phase 1: products (tile starts are multiples of `tile`, itself a multiple of 4)
    for (int64_t b = lo + 4 * (int64_t)threadIdx.x; b < hi; b += 4 * THREADS) {
      const int4 c = ld_stream(reinterpret_cast<const int4 *>(ci + b));
      const Quad<T> q = load_quad(val, b);
      const T x0 = __ldg(x + c.x), x1 = __ldg(x + c.y), x2 = __ldg(x + c.z), x3 = __ldg(x + c.w);
      T *o = prod + (b - lo);
      o[0] = q.v[0] * x0; // entries past hi (padding) are computed but never summed
      o[1] = q.v[1] * x1;
      o[2] = q.v[2] * x2;
      o[3] = q.v[3] * x3;
    }
    if (threadIdx.x == 0) carry_pending[par ^ 1] = -1; // slot this tile may fill
    __syncthreads();

    // phase 2: row sums
    int64_t ra = (t == 0) ? 0 : (int64_t)tile_row[t];
    int64_t rb = tile_row[t + 1];
    if (rb > m - 1) rb = m - 1;
    const int64_t nrows = rb - ra + 1;
    int vec = 1;
    {
      const int64_t mean = (hi - lo) / (nrows > 0 ? nrows : 1);
      while (vec < 32 && vec * 4 < mean) vec <<= 1;
    }
    const int sub = threadIdx.x & (vec - 1);
    const int n_groups = THREADS / vec;
    // warp-uniform trip count: the shuffles below name the whole warp
    for (int64_t rbase = ra + (threadIdx.x & ~31) / vec; rbase <= rb; rbase += n_groups) {
      const int64_t r = rbase + (threadIdx.x & 31) / vec;
      const bool active = r <= rb;
      int64_t rs = 0, re = 0;
      if (active) {
        rs = rp[r];
        re = rp[r + 1];
      }
      const int64_t s = rs > lo ? rs : lo, e = re < hi ? re : hi;
      T acc = 0;
      for (int64_t j = s + sub; j < e; j += vec) acc += prod[j - lo];
      for (int o = vec >> 1; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, vec);
      if (active && sub == 0) {
        if (rs < lo && t > t0) acc += carry[par]; // began in an earlier tile of this CTA
        if (re > hi) {                            // continues: hand over
          carry[par ^ 1] = acc;
          carry_pending[par ^ 1] = (int)r;
        } else if (re > lo || t == 0) {
          y[r] = acc;
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int par = (int)((t1 - t0) & 1);
    const int r = (t1 > t0) ? carry_pending[par] : -1;
    carry_row[blockIdx.x] = r;
    carry_val[blockIdx.x] = (r >= 0) ? carry[par] : (T)0;
  }
}

// adds the cross-CTA carries in CTA order (one thread: the order is part of the result)
template <typename T>
__global__ void merge_fixup_kernel(int n_ctas, const int *__restrict__ carry_row,
                                   const T *__restrict__ carry_val, T *__restrict__ y) {
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int b = 0; b < n_ctas; ++b) {
      const int r = carry_row[b];
      if (r >= 0) y[r] += carry_val[b];
    }
}

} // namespace hspmv
