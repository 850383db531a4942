// sm_100a SpMV kernels.  All HBM-bound: no tensor cores (SpMV is not a dense contraction).
//
//   csr_vector_kernel   row-length-binned CSR: VEC lanes per row (1..32), 4 nonzeros per lane
//                       per step through 128-bit loads, warp-shuffle reduction.
//                       Replaces cuda_spmv, cuda-spmv-csr/spmv.cu:117-182 (thread per row).
//   csr_longrow_kernel  rows of the last bin cut into segments, one CTA each + ordered reduce.
//   csrk_tma_kernel     CSR-k: super-super-row -> CTA, super-row -> warp, row -> VEC lanes.
//                       Each warp owns a private ring of shared-memory stages; one elected lane
//                       brings the next super-row's row_ptr / col / val segments in with TMA
//                       bulk copies (cp.async.bulk + mbarrier complete_tx) while the warp
//                       consumes the current one.  No CTA-wide barrier in the main loop.
//                       Replaces cuSpMV_2 / cuSpMV_3 / cuSpMV_3_vec,
//                       cuda-spmv-csrk/cuda/csrk.cu:320-391 / 244-319 / 184-242.
//                       With k == 1 the same engine runs plain CSR on fixed-row tiles.
//   csr_merge_kernel    nnz-balanced split for power-law matrices (see merge.cuh).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hspmv {

// ------------------------------------------------------------------------------------------
// small PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0,
// both addresses 16 B aligned).  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes,
                                         uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

// Metadata loads of the TMA kernels.  `volatile` + "memory" pins them in program order between the
// other asm statements (the TMA issue before, the mbarrier wait after), so that they are in flight
// while the warp consumes stages.  The metadata queue is two tiles deep: with a single tile of
// slack the load -> TMA-issue dependency sat on the critical path, and a handful of extra
// instructions in front of the load cost 10-17 % of the kernel (profiles/r1_ab_control_path.txt).
__device__ __forceinline__ int ld_meta(const int *p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_meta(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// streaming 128-bit loads of the matrix arrays: read once, keep them out of L1
__device__ __forceinline__ int4 ld_stream(const int4 *p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ld_stream(const double2 *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
               : "=d"(r.x), "=d"(r.y)
               : "l"(p));
  return r;
}

// four consecutive values starting at a 4-aligned element index
template <typename T> struct Quad {
  T v[4];
};
__device__ __forceinline__ Quad<float> load_quad(const float *base, int64_t idx) {
  float4 t = ld_stream(reinterpret_cast<const float4 *>(base + idx));
  Quad<float> q;
  q.v[0] = t.x; q.v[1] = t.y; q.v[2] = t.z; q.v[3] = t.w;
  return q;
}
__device__ __forceinline__ Quad<double> load_quad(const double *base, int64_t idx) {
  double2 a = ld_stream(reinterpret_cast<const double2 *>(base + idx));
  double2 b = ld_stream(reinterpret_cast<const double2 *>(base + idx + 2));
  Quad<double> q;
  q.v[0] = a.x; q.v[1] = a.y; q.v[2] = b.x; q.v[3] = b.y;
  return q;
}

// x gathers.  COH = false: read-only path (ld.global.nc), for x that nobody writes while the kernel
// runs.  COH = true: the tile reads the x halo that a neighbour GPU wrote over NVLink DURING this
// launch (hspmv_spmv_halo); ld.global.nc is outside the memory model there, and an interior tile
// may already have pulled the sector that straddles the own/halo boundary into L1, so these
// gathers go to L2 (ld.global.cg), the point of coherence for peer writes, ordered behind the
// ld.acquire.sys of halo_spin_wait by the __syncwarp that follows it.
template <bool COH, typename T> __device__ __forceinline__ T ld_x(const T *p) {
  return COH ? __ldcg(p) : __ldg(p);
}

template <typename T, int VEC> __device__ __forceinline__ T group_sum(T v) {
#pragma unroll
  for (int o = VEC >> 1; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, VEC);
  return v;
}

// ------------------------------------------------------------------------------------------
// CSR, VEC lanes per row, 128-bit loads
// ------------------------------------------------------------------------------------------
// rows: either the contiguous range [row_base, row_base + n_rows) or row_list[0..n_rows).
// The column / value arrays are padded by >= 4 entries past nnz, so the aligned quads that
// straddle a row's ends are always in bounds; out-of-row lanes are masked by select, never
// multiplied by zero (x may hold Inf/NaN in entries the row does not reference).
template <typename T, int VEC, bool LIST>
__global__ void __launch_bounds__(256)
csr_vector_kernel(int64_t n_rows, const int *__restrict__ row_list, int64_t row_base,
                  const int *__restrict__ rp, const int *__restrict__ ci,
                  const T *__restrict__ val, const T *__restrict__ x, T *__restrict__ y) {
  const int lane = threadIdx.x & (VEC - 1);
  const int64_t n_groups = (int64_t)gridDim.x * blockDim.x / VEC;
  // warp-uniform trip count: every lane of a warp stays in the loop until the warp's first
  // group runs out of rows, so the full-mask shuffles below are always convergent
  for (int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / VEC; g0 < n_rows;
       g0 += n_groups) {
    const int64_t g = g0 + (threadIdx.x & 31) / VEC;
    const bool active = g < n_rows;
    int64_t r = 0, e0 = 0, e1 = 0;
    if (active) {
      r = LIST ? (int64_t)row_list[g] : row_base + g;
      e0 = rp[r];
      e1 = rp[r + 1];
    }
    T acc = 0;
    for (int64_t b = (e0 & ~(int64_t)3) + 4 * lane; b < e1; b += 4 * VEC) {
      const int4 c = ld_stream(reinterpret_cast<const int4 *>(ci + b));
      const Quad<T> q = load_quad(val, b);
      const T x0 = __ldg(x + c.x), x1 = __ldg(x + c.y), x2 = __ldg(x + c.z), x3 = __ldg(x + c.w);
      if (b >= e0 && b + 3 < e1) {
        acc += q.v[0] * x0;
        acc += q.v[1] * x1;
        acc += q.v[2] * x2;
        acc += q.v[3] * x3;
      } else {
        if (b >= e0 && b < e1) acc += q.v[0] * x0;
        if (b + 1 >= e0 && b + 1 < e1) acc += q.v[1] * x1;
        if (b + 2 >= e0 && b + 2 < e1) acc += q.v[2] * x2;
        if (b + 3 >= e0 && b + 3 < e1) acc += q.v[3] * x3;
      }
    }
    acc = group_sum<T, VEC>(acc);
    if (active && lane == 0) y[r] = acc;
  }
}

// Very long rows (a power-law hub holds 10^5 entries): one CTA per SEGMENT of a row, so that one
// row does not serialise on one CTA; seg[3*s..] = {row slot, first entry, end entry}.  The partial
// sums are added per row in segment order by longrow_reduce_kernel (deterministic, no atomics).
template <typename T>
__global__ void __launch_bounds__(256)
csr_longrow_kernel(int64_t n_seg, const int *__restrict__ seg, const int *__restrict__ ci,
                   const T *__restrict__ val, const T *__restrict__ x, T *__restrict__ partial_out) {
  __shared__ T partial[8];
  for (int64_t g = blockIdx.x; g < n_seg; g += gridDim.x) {
    const int64_t e0 = seg[3 * g + 1], e1 = seg[3 * g + 2];
    T acc = 0;
    for (int64_t b = (e0 & ~(int64_t)3) + 4 * (int64_t)threadIdx.x; b < e1; b += 4 * 256) {
      const int4 c = ld_stream(reinterpret_cast<const int4 *>(ci + b));
      const Quad<T> q = load_quad(val, b);
      if (b >= e0 && b < e1) acc += q.v[0] * __ldg(x + c.x);
      if (b + 1 >= e0 && b + 1 < e1) acc += q.v[1] * __ldg(x + c.y);
      if (b + 2 >= e0 && b + 2 < e1) acc += q.v[2] * __ldg(x + c.z);
      if (b + 3 >= e0 && b + 3 < e1) acc += q.v[3] * __ldg(x + c.w);
    }
    acc = group_sum<T, 32>(acc);
    if ((threadIdx.x & 31) == 0) partial[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      T s = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += partial[w];
      partial_out[g] = s;
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void longrow_reduce_kernel(int64_t n_rows, const int *__restrict__ row_list, const int *__restrict__ seg_ptr,
                                      const T *__restrict__ partial, T *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    T s = 0;
    for (int g = seg_ptr[i]; g < seg_ptr[i + 1]; ++g) s += partial[g];
    y[row_list[i]] = s;
  }
}

// rows with no stored entry: y = 0 (only launched when the matrix has such rows and the
// main kernel works from row lists)
template <typename T>
__global__ void zero_rows_kernel(int64_t n_rows, const int *__restrict__ row_list,
                                 T *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows;
       i += (int64_t)gridDim.x * blockDim.x)
    y[row_list[i]] = 0;
}

// ------------------------------------------------------------------------------------------
// CSR_BLOCK: row blocks sized by nonzero count, products staged in shared memory
// ------------------------------------------------------------------------------------------
// For matrices that live in L2 (a launch lasts tens of microseconds: the persistent TMA ring's
// prologue and its one-tile-per-warp latency chain are most of that) and for ragged rows (no
// padding of every row to a common slot count).  blk_row[b] .. blk_row[b+1] is a block of
// consecutive rows with at most `cap` nonzeros (host_index.h: block_rows), one CTA each, grid =
// number of blocks, many CTAs per SM -- latency is hidden by occupancy, not by a pipeline:
//   phase 1  the block's nonzero range is read with coalesced 128-bit loads, every thread keeps
//            4 x gathers in flight per step, the products go to shared memory;
//   phase 2  VEC lanes per row add the row's products from shared memory (stride VEC), warp
//            shuffle finishes the sum.
// A row longer than `cap` is a block of its own and never touches shared memory for products:
// every thread adds its strided share in a register, then one block reduction.
// The aligned quads that straddle the ends of the block's range multiply entries of the
// neighbouring rows (or the zero pad) by valid x entries; those products are never added.
template <typename T, int VEC, int THREADS>
__global__ void __launch_bounds__(THREADS)
csr_block_kernel(int n_blocks, int cap, const int *__restrict__ blk_row, const int *__restrict__ rp,
                 const int *__restrict__ ci, const T *__restrict__ val, const T *__restrict__ x, T *__restrict__ y) {
  extern __shared__ __align__(16) unsigned char block_smem[];
  T *prod = reinterpret_cast<T *>(block_smem);                 // cap + 8 products
  int *s_rp = reinterpret_cast<int *>(prod + cap + 8);         // row pointers of the block
  __shared__ T warp_part[THREADS / 32];
  const int b = blockIdx.x;
  const int r0 = blk_row[b], r1 = blk_row[b + 1];
  const int rows = r1 - r0;
  const int64_t e0 = rp[r0], e1 = rp[r1];
  const int64_t ea = e0 & ~(int64_t)3;
  if (rows == 1 && e1 - ea > cap) { // one long row: register accumulation + block reduction
    T acc = 0;
    for (int64_t j = ea + 4 * (int64_t)threadIdx.x; j < e1; j += 4 * THREADS) {
      const int4 c = ld_stream(reinterpret_cast<const int4 *>(ci + j));
      const Quad<T> q = load_quad(val, j);
      const T x0 = __ldg(x + c.x), x1 = __ldg(x + c.y), x2 = __ldg(x + c.z), x3 = __ldg(x + c.w);
      if (j >= e0 && j + 3 < e1) {
        acc += q.v[0] * x0;
        acc += q.v[1] * x1;
        acc += q.v[2] * x2;
        acc += q.v[3] * x3;
      } else {
        if (j >= e0 && j < e1) acc += q.v[0] * x0;
        if (j + 1 >= e0 && j + 1 < e1) acc += q.v[1] * x1;
        if (j + 2 >= e0 && j + 2 < e1) acc += q.v[2] * x2;
        if (j + 3 >= e0 && j + 3 < e1) acc += q.v[3] * x3;
      }
    }
    acc = group_sum<T, 32>(acc);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      T s = 0;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) s += warp_part[w];
      y[r0] = s;
    }
    return;
  }
  for (int i = threadIdx.x; i <= rows; i += THREADS) s_rp[i] = rp[r0 + i] - (int)ea;
  // phase 1
  for (int64_t j = ea + 4 * (int64_t)threadIdx.x; j < e1; j += 4 * THREADS) {
    const int4 c = ld_stream(reinterpret_cast<const int4 *>(ci + j));
    const Quad<T> q = load_quad(val, j);
    const T x0 = __ldg(x + c.x), x1 = __ldg(x + c.y), x2 = __ldg(x + c.z), x3 = __ldg(x + c.w);
    T *o = prod + (j - ea);
    o[0] = q.v[0] * x0;
    o[1] = q.v[1] * x1;
    o[2] = q.v[2] * x2;
    o[3] = q.v[3] * x3;
  }
  __syncthreads();
  // phase 2 (warp-uniform trip count: the shuffles name the whole warp)
  const int sub = threadIdx.x & (VEC - 1);
  for (int rb = (threadIdx.x & ~31) / VEC; rb < rows; rb += THREADS / VEC) {
    const int r = rb + (threadIdx.x & 31) / VEC;
    int rs = 0, re = 0;
    if (r < rows) {
      rs = s_rp[r];
      re = s_rp[r + 1];
    }
    T acc = 0;
    for (int j = rs + sub; j < re; j += VEC) acc += prod[j];
    acc = group_sum<T, VEC>(acc);
    if (r < rows && sub == 0) y[r0 + r] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// CSR-k, TMA-staged super-rows
// ------------------------------------------------------------------------------------------
template <typename T> struct CsrkParams {
  int k;              // 1: implicit tiles of rows_per_tile rows, 2: sr_ptr, 3: ssr_ptr + sr_ptr
  int64_t m;          // rows
  // k == 1: up to three row ranges, tiled one after the other (hspmv_spmv_range passes one,
  // hspmv_spmv_halo passes interior | low boundary | high boundary); tiles_before[i] = number
  // of tiles in ranges < i, tiles_before[n_ranges] = n_sr
  int n_ranges;
  int64_t range_lo[3], range_hi[3], tiles_before[4];
  // fused halo wait (hspmv_spmv_halo): the tiles of ranges 1 and 2 read x entries that a
  // neighbour GPU writes over NVLink; before a warp consumes its first such tile it spins
  // (bounded) until every wait_flags[i] >= wait_step.  n_wait == 0: no waiting.
  const int *wait_flags;
  int n_wait, wait_step;
  int *timed_out;
  int64_t n_sr;       // super-rows (k == 1: number of tiles)
  int64_t n_ssr;      // super-super-rows (k == 3)
  const uint32_t *ssr_ptr; // [n_ssr+1]
  const uint32_t *sr_ptr;  // [n_sr+1]
  const uint32_t *sr_e;    // [n_sr+1] = rp[sr_ptr[s]]: device-only helper, saves one dependent load
  const int *rp;
  const int *ci;
  const T *val;
  const T *x;
  T *y;
  int rows_per_tile;  // k == 1
  int stages;         // ring depth per warp
  int cap_nnz;        // stage capacity in nonzeros (multiple of 4)
  int cap_rows;       // stage capacity in rows (multiple of 4)
};

// sequence of super-rows a warp processes
struct SrCursor {
  int64_t cur, end, step;     // k < 3: cur, cur+step, ... < end
  int64_t s, s_step, n_ssr;   // k == 3: CTA walks super-super-rows s, s+s_step, ...
  const uint32_t *ssr_ptr;
  uint32_t pre_lo, pre_hi;    // k == 3: super-row range of the NEXT super-super-row, loaded one
                              // switch ahead so that its latency is off the warp's critical path
  int carry;                  // k == 3: super-rows this CTA has dealt so far, mod W.  The CTA's
                              // super-rows are dealt round-robin ACROSS super-super-rows, so a
                              // 17-super-row SSR does not always give its extra one to warp 0
  int w, W;
  bool three;
  __device__ __forceinline__ void prefetch_range(int64_t ss) {
    if (ss < n_ssr) {
      pre_lo = ld_meta(ssr_ptr + ss);
      pre_hi = ld_meta(ssr_ptr + ss + 1);
    }
  }
  __device__ __forceinline__ bool next(int64_t &sr) {
    if (!three) {
      if (cur >= end) return false;
      sr = cur;
      cur += step;
      return true;
    }
    for (;;) {
      if (cur < end) {
        sr = cur;
        cur += W;
        return true;
      }
      s += s_step;
      if (s >= n_ssr) return false;
      int first = w - carry;
      if (first < 0) first += W;
      cur = (int64_t)pre_lo + first;
      end = (int64_t)pre_hi;
      carry = (int)((carry + (pre_hi - pre_lo)) % (uint32_t)W);
      prefetch_range(s + s_step);
    }
  }
};

struct SrMeta {
  int r0, r1;     // rows
  int e0, e1;     // nonzeros
};

__host__ __device__ inline size_t csrk_stage_bytes(int cap_nnz, int cap_rows, size_t vsize) {
  // meta (16 B) | row_ptr window | col window | val window ; every piece a multiple of 16 B
  return 16 + (size_t)(cap_rows + 8) * 4 + (size_t)(cap_nnz + 8) * 4 + (size_t)(cap_nnz + 8) * vsize;
}

// Rows [r0,r1) of one super-row: row -> group of VEC lanes (VEC = 1: the CSR-k "row -> lane").
// A lane's share of a row (every VEC-th nonzero) is consumed in two loops:
//   MAIN  unmasked chunks of MAIN slots: all MAIN column reads, value reads and x gathers are
//         issued before the first FMA, so a lane keeps MAIN independent gathers in flight (the
//         gathers, not the streamed matrix data, are what this loop waits for -- see
//         profiles/r1_*): 5 instructions per nonzero;
//   TAIL  what is left (< MAIN slots) in masked chunks of TAIL slots: a slot past the row's end
//         re-reads the row's LAST entry with value 0, so the gather stays unconditional and inside
//         the row's own column set (0 * x[c] only meets an x entry this row references anyway).
// MAIN = 0 disables the unmasked loop.  Index type I: int for the staged path (offsets into the
// shared-memory window), int64_t for the unstaged path (global nonzero indices).
template <typename T, int VEC, int MAIN, int TAIL, typename I, bool COH = false>
__device__ __forceinline__ void csrk_rows(int r0, int r1, const int *__restrict__ rp_w, int rbase,
                                          const int *__restrict__ ci_w, const T *__restrict__ val_w,
                                          I ebase, const T *__restrict__ x, T *__restrict__ y,
                                          int lane) {
  const int sub = lane & (VEC - 1);
  const int gid = lane / VEC;
  // warp-uniform trip count (full-mask shuffles inside)
  for (int rb = r0; rb < r1; rb += 32 / VEC) {
    const int r = rb + gid;
    const bool active = r < r1;
    I rs = 0, re = 0;
    if (active) {
      rs = (I)rp_w[r - rbase] - ebase;
      re = (I)rp_w[r + 1 - rbase] - ebase;
    }
    T acc = 0;
    I j = rs + sub;
    if (MAIN > 0) {
      for (; j + (I)((MAIN - 1) * VEC) < re; j += VEC * MAIN) {
        int c[MAIN > 0 ? MAIN : 1];
        T v[MAIN > 0 ? MAIN : 1], xv[MAIN > 0 ? MAIN : 1];
#pragma unroll
        for (int u = 0; u < MAIN; ++u) {
          c[u] = ci_w[j + u * VEC];
          v[u] = val_w[j + u * VEC];
        }
#pragma unroll
        for (int u = 0; u < MAIN; ++u) xv[u] = ld_x<COH>(x + c[u]);
#pragma unroll
        for (int u = 0; u < MAIN; ++u) acc += v[u] * xv[u];
      }
    }
    const I last = re - 1;
    for (; j < re; j += VEC * TAIL) {
      int c[TAIL];
      T v[TAIL], xv[TAIL];
#pragma unroll
      for (int u = 0; u < TAIL; ++u) {
        const I jj = j + u * VEC;
        const I jc = jj < last ? jj : last;
        c[u] = ci_w[jc];
        v[u] = val_w[jc];
        if (jj > last) v[u] = (T)0;
      }
#pragma unroll
      for (int u = 0; u < TAIL; ++u) xv[u] = ld_x<COH>(x + c[u]);
#pragma unroll
      for (int u = 0; u < TAIL; ++u) acc += v[u] * xv[u];
    }
    acc = group_sum<T, VEC>(acc);
    if (active && sub == 0) y[r] = acc;
  }
}

// First boundary tile of a warp in hspmv_spmv_halo: the halo part of x must have arrived.  The
// tile's matrix data is already staged (it is local); only the gathers after this need the halo.
// Kept out of line: it runs once per warp per launch and must not perturb the main loop's code.
static __device__ __noinline__ void halo_spin_wait(const int *flags, int n_wait, int step, int *timed_out, int lane) {
  const long long t0 = clock64();
  if (lane < n_wait) {
    for (;;) {
      int v;
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + lane) : "memory");
      if (v >= step) break;
      if (clock64() - t0 > 4000000000LL) {
        atomicExch(timed_out, 1);
        break;
      }
      __nanosleep(100);
    }
  }
  __syncwarp();
  // diagnostics next to the timeout word: [1] longest spin of any warp so far (in 1024-cycle
  // units), [2] number of warp-level waits, [3] total spin (1024-cycle units), [4..9] histogram
  // of the spins: < 8, < 32, < 128, < 512, < 2048, >= 2048 units
  if (lane == 0) {
    const int kc = (int)((clock64() - t0) >> 10);
    atomicMax(timed_out + 1, kc);
    atomicAdd(timed_out + 2, 1);
    atomicAdd(timed_out + 3, kc);
    const int b = kc < 8 ? 0 : kc < 32 ? 1 : kc < 128 ? 2 : kc < 512 ? 3 : kc < 2048 ? 4 : 5;
    atomicAdd(timed_out + 4 + b, 1);
  }
}

// Halo tiles (a few per launch, hspmv_spmv_halo only) gather x through L2.  Kept out of line, so that the code the
// other 98 % of the tiles run -- and its register allocation -- stays what it is without the halo support.
template <typename T, int VEC, int MAIN, int TAIL>
static __device__ __noinline__ void csrk_rows_halo_staged(int r0, int r1, const int *rp_w, const int *ci_w, const T *va_w, int e0,
                                                          const T *x, T *y, int lane) {
  csrk_rows<T, VEC, MAIN, TAIL, int, true>(r0, r1, rp_w, r0 & ~3, ci_w, va_w, e0 & ~3, x, y, lane);
}
template <typename T, int VEC, int MAIN, int TAIL>
static __device__ __noinline__ void csrk_rows_halo_direct(int r0, int r1, const int *rp, const int *ci, const T *val, const T *x, T *y,
                                                          int lane) {
  csrk_rows<T, VEC, MAIN, TAIL, int64_t, true>(r0, r1, rp, 0, ci, val, (int64_t)0, x, y, lane);
}

// HALO = false: the kernel every single-GPU launch runs; it contains no trace of the halo support.  HALO = true
// (hspmv_spmv_halo): up to three row ranges, the in-kernel flag wait and L2 gathers for the halo tiles.
// WIDE = true: compiled for 640 threads per CTA (96 registers per thread instead of 128), launched with 18 warps.  Only
// instantiated for the shapes that fit 96 registers without spilling and were measured faster that way (the 27-point
// stencil shapes: 3.136 -> 3.082 ms on config 5; 20 warps are slower again -- their stages leave 20 KB of L1 for the x
// gathers -- and the masked shapes spill; profiles/r2_ab_warps_per_sm.txt).
template <typename T, int VEC, int MAIN, int TAIL, bool HALO, bool WIDE = false>
__global__ void __launch_bounds__(WIDE ? 640 : 512, 1) csrk_tma_kernel(const CsrkParams<T> p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const int W = blockDim.x >> 5;
  const int S = p.stages;
  const size_t stage_bytes = csrk_stage_bytes(p.cap_nnz, p.cap_rows, sizeof(T));
  // layout: [W*S mbarriers][warp 0: S stages][warp 1: S stages]...
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw) + (size_t)w * S;
  unsigned char *ring = smem_raw + (((size_t)W * S * 8 + 127) & ~(size_t)127) + (size_t)w * S * stage_bytes;

  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(smem_u32(bars + s), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncwarp();

  SrCursor cur;
  cur.three = (p.k == 3);
  cur.w = w;
  cur.W = W;
  cur.ssr_ptr = p.ssr_ptr;
  cur.n_ssr = p.n_ssr;
  if (cur.three) {
    cur.s = (int64_t)blockIdx.x - gridDim.x; // first next() advances to blockIdx.x
    cur.s_step = gridDim.x;
    cur.cur = 0;
    cur.end = 0;
    cur.step = 0;
    cur.pre_lo = cur.pre_hi = 0;
    cur.carry = 0;
    cur.prefetch_range(blockIdx.x);
  } else {
    cur.cur = (int64_t)blockIdx.x * W + w;
    cur.step = (int64_t)gridDim.x * W;
    cur.end = p.n_sr;
    cur.s = cur.s_step = 0;
    cur.carry = 0;
  }

  const uint64_t policy = l2_policy_evict_first();

  auto load_meta = [&](int64_t sr) -> SrMeta {
    SrMeta mt;
    if (p.k == 1) {
      int64_t r0, hi;
      if (!HALO || sr < p.tiles_before[1]) { // (warp-uniform branches; a single range always takes the first)
        r0 = p.range_lo[0] + sr * p.rows_per_tile;
        hi = p.range_hi[0];
      } else if (sr < p.tiles_before[2]) {
        r0 = p.range_lo[1] + (sr - p.tiles_before[1]) * p.rows_per_tile;
        hi = p.range_hi[1];
      } else {
        r0 = p.range_lo[2] + (sr - p.tiles_before[2]) * p.rows_per_tile;
        hi = p.range_hi[2];
      }
      const int64_t r1 = (r0 + p.rows_per_tile < hi) ? r0 + p.rows_per_tile : hi;
      mt.r0 = (int)r0;
      mt.r1 = (int)r1;
      mt.e0 = ld_meta(p.rp + r0);
      mt.e1 = ld_meta(p.rp + r1);
    } else {
      mt.r0 = (int)ld_meta(p.sr_ptr + sr);
      mt.r1 = (int)ld_meta(p.sr_ptr + sr + 1);
      mt.e0 = (int)ld_meta(p.sr_e + sr);
      mt.e1 = (int)ld_meta(p.sr_e + sr + 1);
    }
    return mt;
  };

  // post the TMA copies (or an empty arrival for an oversized super-row) for `mt` into stage st
  auto issue = [&](const SrMeta &mt, int st) {
    if (lane == 0) {
      unsigned char *base = ring + (size_t)st * stage_bytes;
      const int rows = mt.r1 - mt.r0, cnt = mt.e1 - mt.e0;
      const bool staged = rows <= p.cap_rows && cnt <= p.cap_nnz && cnt > 0;
      int *meta = reinterpret_cast<int *>(base);
      meta[0] = mt.r0;
      meta[1] = mt.r1;
      meta[2] = mt.e0;
      meta[3] = staged ? mt.e1 : -mt.e1 - 1; // negative: not staged, read from global
      const uint32_t bar = smem_u32(bars + st);
      if (staged) {
        const int rb = mt.r0 & ~3, rt = (mt.r1 + 1 + 3) & ~3; // row_ptr entries r0..r1
        const int64_t eb = mt.e0 & ~3, et = ((int64_t)mt.e1 + 3) & ~(int64_t)3;
        const uint32_t rp_bytes = (uint32_t)(rt - rb) * 4u;
        const uint32_t ci_bytes = (uint32_t)(et - eb) * 4u;
        const uint32_t va_bytes = (uint32_t)(et - eb) * (uint32_t)sizeof(T);
        unsigned char *rp_w = base + 16;
        unsigned char *ci_w = rp_w + (size_t)(p.cap_rows + 8) * 4;
        unsigned char *va_w = ci_w + (size_t)(p.cap_nnz + 8) * 4;
        // No proxy fence here: the stage was only READ through the generic proxy, every one of
        // those reads has returned (their values fed the FMAs whose sums were stored before the
        // __syncwarp that precedes this call), and a write-after-read hand-over to the async
        // proxy needs execution ordering only -- the same contract as a TMA pipeline's
        // consumer-release / producer-acquire pair.  (A fence.proxy.async costs a MEMBAR.ALL.CTA
        // that also waits for the y stores; it was 8 % of the kernel, profiles/r1_ab_control_path.txt.)
        mbar_arrive_expect_tx(bar, rp_bytes + ci_bytes + va_bytes);
        bulk_g2s(smem_u32(rp_w), p.rp + rb, rp_bytes, bar, policy);
        bulk_g2s(smem_u32(ci_w), p.ci + eb, ci_bytes, bar, policy);
        bulk_g2s(smem_u32(va_w), p.val + eb, va_bytes, bar, policy);
      } else {
        mbar_arrive(bar);
      }
    }
  };

  // ---- prologue: fill the ring, then keep the metadata of the next TWO tiles in registers
  int64_t sr;
  int filled = 0;
  for (; filled < S; ++filled) {
    if (!cur.next(sr)) break;
    issue(load_meta(sr), filled);
  }
  SrMeta mq0, mq1; // mq0: next tile to issue, mq1: the one after
  mq0.r0 = mq0.r1 = mq0.e0 = mq0.e1 = 0;
  mq1 = mq0;
  bool have0 = cur.next(sr);
  if (have0) mq0 = load_meta(sr);
  bool have1 = have0 && cur.next(sr);
  if (have1) mq1 = load_meta(sr);

  int live = filled;
  // re-arm stage `stg` with the next tile of this warp and pull one more tile's metadata
  auto advance = [&](int stg) {
    --live;
    if (have0) {
      issue(mq0, stg);
      ++live;
      mq0 = mq1;
      have0 = have1;
      have1 = have1 && cur.next(sr);
      if (have1) mq1 = load_meta(sr);
    }
  };

  // ---- main loop
  int st = 0;
  uint32_t parity = 0;
  // Tiles come to a warp in increasing index order, interior range first: the halo wait sits in
  // front of the (wait_at)-th tile this warp consumes; -1 = never.
  int consumed = 0, wait_at = -1;
  if (HALO && p.k == 1 && p.n_wait > 0) {
    const int64_t gw = (int64_t)blockIdx.x * W + w, GW = (int64_t)gridDim.x * W;
    const int64_t n_int = p.tiles_before[1] > gw ? (p.tiles_before[1] - gw + GW - 1) / GW : 0;
    wait_at = (int)n_int;
  }
  constexpr int NS = MAIN + TAIL; // register slots per lane in the single-pass path
  const int sub = lane & (VEC - 1);
  const int gid = lane / VEC;
  while (live > 0) {
    const uint32_t bar = smem_u32(bars + st);
    mbar_wait(bar, parity);
    unsigned char *base = ring + (size_t)st * stage_bytes;
    const int *meta = reinterpret_cast<const int *>(base);
    const int r0 = meta[0], r1 = meta[1], e0 = meta[2], e1m = meta[3];
    if (HALO && consumed == wait_at) halo_spin_wait(p.wait_flags, p.n_wait, p.wait_step, p.timed_out, lane);
    const bool halo_tile = HALO && wait_at >= 0 && consumed >= wait_at; // warp-uniform: this tile gathers from the halo
    if (HALO) ++consumed;
    bool released = false;
    if (e1m >= 0) {
      const int *rp_w = reinterpret_cast<const int *>(base + 16);
      const int *ci_w = rp_w + (p.cap_rows + 8);
      const T *va_w = reinterpret_cast<const T *>(reinterpret_cast<const unsigned char *>(ci_w) +
                                                   (size_t)(p.cap_nnz + 8) * 4);
      // Single-pass path: the tile has at most one row per lane group.  The first VEC * NS entries
      // of every row are read into registers FIRST, the stage is handed back to TMA at once (the
      // next tile's copies fly during the x gathers instead of after them), and the gathers and
      // FMAs run from registers.  The rare row that is longer has its excess consumed straight
      // from the stage before the release (one extra round of gathers for that tile only).
      // (halo tiles -- a few per launch -- take the csrk_rows<COH> path below, so that this path keeps one gather flavour)
      const bool single = (r1 - r0) <= 32 / VEC && !halo_tile;
      int rs = 0, re = 0;
      const int r = r0 + gid;
      const bool active = r < r1;
      if (single) {
        if (active) {
          const int rbase = r0 & ~3, ebase = e0 & ~3;
          rs = rp_w[r - rbase] - ebase;
          re = rp_w[r + 1 - rbase] - ebase;
        }
        int c[NS];
        T v[NS];
        const int j0 = rs + sub, last = re - 1;
        T acc = 0;
        if (!__all_sync(0xffffffffu, re - rs <= VEC * NS)) {
          for (int j = j0 + NS * VEC; __any_sync(0xffffffffu, j < re); j += 4 * VEC) {
            int cx[4];
            T vx[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int jj = j + u * VEC;
              const bool in = jj < re;
              cx[u] = ci_w[in ? jj : 0];
              vx[u] = in ? va_w[jj] : (T)0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const T xg = __ldg(p.x + cx[u]);
              if (j + u * VEC < re) acc += vx[u] * xg; // predicated: an idle slot must not meet an inf in x
            }
          }
        }
        if (MAIN > 0 && j0 + (MAIN - 1) * VEC < re) {
#pragma unroll
          for (int u = 0; u < MAIN; ++u) {
            c[u] = ci_w[j0 + u * VEC];
            v[u] = va_w[j0 + u * VEC];
          }
#pragma unroll
          for (int u = MAIN; u < NS; ++u) { // the few slots past the unmasked chunk
            const int j = j0 + u * VEC;
            const int jc = j < last ? j : last;
            c[u] = ci_w[jc];
            v[u] = va_w[jc];
            if (j > last) v[u] = (T)0;
          }
        } else {
#pragma unroll
          for (int u = 0; u < NS; ++u) {
            c[u] = 0;
            v[u] = (T)0;
          }
          if (re > rs) {
#pragma unroll
            for (int u = 0; u < NS; ++u) {
              const int j = j0 + u * VEC;
              const int jc = j < last ? j : last; // slots past the end re-read the row's last entry with value 0
              c[u] = ci_w[jc];
              const T t = va_w[jc];
              v[u] = j > last ? (T)0 : t;
            }
          }
        }
        __syncwarp(); // every lane has its share of the stage in registers
        advance(st);
        released = true;
        T xv[NS];
#pragma unroll
        for (int u = 0; u < NS; ++u) xv[u] = __ldg(p.x + c[u]);
#pragma unroll
        for (int u = 0; u < NS; ++u) acc += v[u] * xv[u]; // same order as csrk_rows: a row's sum does not depend on the tiling
        if (re <= rs) acc = (T)0; // empty row: its idle slots multiplied 0 by x[0], which may be inf
        acc = group_sum<T, VEC>(acc);
        if (active && sub == 0) p.y[r] = acc;
      } else if (halo_tile) {
        csrk_rows_halo_staged<T, VEC, MAIN, TAIL>(r0, r1, rp_w, ci_w, va_w, e0, p.x, p.y, lane);
      } else {
        csrk_rows<T, VEC, MAIN, TAIL, int>(r0, r1, rp_w, r0 & ~3, ci_w, va_w, e0 & ~3, p.x, p.y, lane);
      }
    } else if (halo_tile) {
      csrk_rows_halo_direct<T, VEC, MAIN, TAIL>(r0, r1, p.rp, p.ci, p.val, p.x, p.y, lane);
    } else {
      csrk_rows<T, VEC, MAIN, TAIL, int64_t>(r0, r1, p.rp, 0, p.ci, p.val, (int64_t)0, p.x, p.y, lane);
    }
    if (!released) {
      __syncwarp();
      advance(st);
    }
    if (++st == S) {
      st = 0;
      parity ^= 1;
    }
  }
}

// ------------------------------------------------------------------------------------------
// vector helpers
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void gather_kernel(int64_t n, const uint32_t *__restrict__ perm,
                              const T *__restrict__ src, T *__restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[perm[i]];
}
template <typename T>
__global__ void scatter_kernel(int64_t n, const uint32_t *__restrict__ perm,
                               const T *__restrict__ src, T *__restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[perm[i]] = src[i];
}
// hspmv_spmv_timed, flush mode 2: read a second buffer larger than L2 after the written one, so that L2 ends up full of
// CLEAN foreign lines (a kernel timed right after a memset also pays for the write-back of the memset's dirty lines)
static __global__ void flush_read_kernel(const int4 *__restrict__ p, int64_t n16, int *sink) {
  int acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 v = p[i];
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x5bd1e995) *sink = acc; // never true for a zeroed buffer: keeps the loads alive
}

// hspmv_relabel_hub_columns: how often every column is referenced; then ci[j] = new_id[ci[j]]
static __global__ void col_count_kernel(int64_t nnz, const int *__restrict__ ci, unsigned int *__restrict__ count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(count + ci[i], 1u);
}
static __global__ void col_relabel_kernel(int64_t nnz, const uint32_t *__restrict__ new_id, int *__restrict__ ci) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
    ci[i] = (int)new_id[ci[i]];
}

// sr_e[s] = rp[sr_ptr[s]]
static __global__ void sr_offsets_kernel(int64_t n_sr, const uint32_t *__restrict__ sr_ptr,
                                  const int *__restrict__ rp, uint32_t *__restrict__ sr_e) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_sr;
       i += (int64_t)gridDim.x * blockDim.x)
    sr_e[i] = (uint32_t)rp[sr_ptr[i]];
}

// hspmv_slice_rows: dst[i] = src[i] - sub; *bad is raised when a result leaves [0, limit)
static __global__ void rebase_kernel(int64_t n, const int *__restrict__ src, int sub, int limit, int *__restrict__ dst,
                                     int *__restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = src[i] - sub;
    dst[i] = t;
    if (t < 0 || t >= limit) *bad = 1;
  }
}

// Rows whose columns all lie in [own_lo, own_hi) are "interior" (need no halo).  Reports the last
// non-interior row of the first half (+1) and the first non-interior row of the second half, so
// that [row_a, row_b) is a contiguous all-interior range.  One warp per row.
static __global__ void interior_rows_kernel(int64_t m, const int *__restrict__ rp, const int *__restrict__ ci,
                                     int own_lo, int own_hi, int *row_a, int *row_b) {
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = (int64_t)gridDim.x * blockDim.x / 32;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32; r < m; r += n_warps) {
    bool outside = false;
    for (int64_t j = rp[r] + lane; j < rp[r + 1]; j += 32) {
      const int c = ci[j];
      outside = outside || c < own_lo || c >= own_hi;
    }
    outside = __any_sync(0xffffffffu, outside);
    if (outside && lane == 0) {
      if (r < m / 2) atomicMax(row_a, (int)r + 1);
      else atomicMin(row_b, (int)r);
    }
  }
}

static __global__ void col_minmax_kernel(int64_t nnz, const int *__restrict__ ci, int *out_min,
                                  int *out_max) {
  int lo = INT32_MAX, hi = -1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = ci[i];
    lo = min(lo, c);
    hi = max(hi, c);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_down_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_down_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(out_min, lo);
    atomicMax(out_max, hi);
  }
}

} // namespace hspmv
