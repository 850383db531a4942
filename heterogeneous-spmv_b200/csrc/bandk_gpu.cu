// Band-k on the device: the O(nnz) parts of the CSR-k build run in HBM.
//
//   level-1 coarse graph    BAND_k::handCoarsen, spmv-csrk/csrk.cpp:1295-1420
//   A' = P A P^T            CSRk_Graph::reorderA, spmv-csrk/csrk.cpp:548-676
//
// Everything that works on the coarse graph (RCM with its std::sort tie order, level 2,
// uncoarsening) stays in bandk.cpp: it is 16x smaller and sequential by definition.  The results
// are bit-identical to the host builder's, which the GPU tests check array by array.
//
// Coarse graph, per group cv of contiguous rows [start[cv], start[cv+1]):
//   1. one CTA gathers grp[w] for every entry w >= start[cv] of the group's rows into shared
//      memory, sorts them (bitonic), run-length encodes -> the group's UPPER neighbours
//      (same or later group) with multiplicities, parked in a scratch area at the group's own
//      nonzero offset; every strict-upper entry also counts one LOWER entry of its target;
//   2. adjacency offsets = scan(lower + upper counts);
//   3. the upper lists are copied behind the lower part, the mirrored entries are scattered into
//      the targets' lower parts (atomic cursor), and each lower part is then sorted by source id
//      -- the order the reference gets by inserting mirrors in increasing source order
//      (csrk.cpp:1350-1354) and sorting per vertex (:1366-1368).
// Segments longer than the shared-memory capacities raise `overflow`; the caller then uses the
// host builder, so there is no size at which the build fails.
#include "bandk_gpu.h"

#include <cuda_runtime.h>

namespace hspmv {

namespace {

constexpr int GROUP_CAP = 8192;  // entries of one group's rows a CTA can sort (keys 32 KB + heads 32 KB)
constexpr int GROUP_THREADS = 256;
constexpr int WARP_CAP = BANDK_MAX_ROW; // (key, payload) pairs a warp can sort in shared memory
constexpr int SORT_WARPS = 8;

#define BK_CU(call)                        \
  do {                                     \
    cudaError_t e_ = (call);               \
    if (e_ != cudaSuccess) return (int)HSPMV_ECUDA; \
  } while (0)

__device__ __forceinline__ int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ---- exclusive scan of u32 (n inputs -> n+1 outputs, out[n] = total) -------------------------
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ u32 block_exclusive(u32 v, u32 *warp_sums, u32 &total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  u32 inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    u32 s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      u32 t = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = s; // inclusive over warps
  }
  __syncthreads();
  total = warp_sums[SCAN_THREADS / 32 - 1];
  const u32 before = w ? warp_sums[w - 1] : 0;
  __syncthreads();
  return before + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(const u32 *in, int64_t n, u32 *sums) {
  __shared__ u32 ws[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
  u32 s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += in[base + i];
  u32 total;
  block_exclusive(s, ws, total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const u32 *in, int64_t n, const u32 *offsets, u32 *out) {
  __shared__ u32 ws[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
  u32 v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = base + i < n ? in[base + i] : 0;
    s += v[i];
  }
  u32 total;
  u32 run = block_exclusive(s, ws, total) + (offsets ? offsets[blockIdx.x] : 0);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i <= n) out[base + i] = run; // out[n] = grand total
    run += v[i];
  }
}

int scan_exclusive(const u32 *in, int64_t n, u32 *out, cudaStream_t st, int64_t &launches) {
  const int64_t blocks = (n + 1 + SCAN_CHUNK - 1) / SCAN_CHUNK;
  if (blocks == 1) {
    scan_apply_kernel<<<1, SCAN_THREADS, 0, st>>>(in, n, nullptr, out);
    ++launches;
    return HSPMV_OK;
  }
  u32 *sums = nullptr, *offs = nullptr;
  BK_CU(cudaMalloc(&sums, (size_t)blocks * 4));
  BK_CU(cudaMalloc(&offs, (size_t)(blocks + 1) * 4));
  scan_sums_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(in, n, sums);
  ++launches;
  int rc = scan_exclusive(sums, blocks, offs, st, launches);
  if (!rc) {
    scan_apply_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(in, n, offs, out);
    ++launches;
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = HSPMV_ECUDA;
  }
  cudaFree(sums);
  cudaFree(offs);
  return rc;
}

// ---- sorting networks ---------------------------------------------------------------------
// 32 (key, payload) pairs, one per lane, ascending by key; pad lanes carry key 0xFFFFFFFF
__device__ __forceinline__ void sort32_pairs(u32 &key, u32 &val, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const u32 ok = __shfl_xor_sync(0xffffffffu, key, j), ov = __shfl_xor_sync(0xffffffffu, val, j);
      const bool up = (lane & k) == 0, lower = (lane & j) == 0;
      const bool take = (lower == up) ? (ok < key) : (ok > key);
      if (take) {
        key = ok;
        val = ov;
      }
    }
}

// n2 (power of two) pairs in shared memory, sorted by one warp
__device__ __forceinline__ void warp_bitonic_pairs(u32 *sk, u32 *sv, int n2, int lane) {
  for (int k = 2; k <= n2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n2; i += 32) {
        const int x = i ^ j;
        if (x > i) {
          const u32 a = sk[i], b = sk[x];
          if ((a > b) == ((i & k) == 0)) {
            sk[i] = b;
            sk[x] = a;
            const u32 t = sv[i];
            sv[i] = sv[x];
            sv[x] = t;
          }
        }
      }
      __syncwarp();
    }
}

// n2 keys in shared memory, sorted by the whole CTA
__device__ __forceinline__ void block_bitonic_keys(u32 *sk, int n2) {
  for (int k = 2; k <= n2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int x = i ^ j;
        if (x > i) {
          const u32 a = sk[i], b = sk[x];
          if ((a > b) == ((i & k) == 0)) {
            sk[i] = b;
            sk[x] = a;
          }
        }
      }
      __syncthreads();
    }
}

// ---- level-1 coarse graph -----------------------------------------------------------------
__global__ void group_of_row_kernel(int64_t n, int64_t nc, const u32 *start, u32 *grp) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = nc; // last group with start <= v
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (start[mid] <= (u32)v)
        lo = mid;
      else
        hi = mid;
    }
    grp[v] = (u32)lo;
  }
}

__global__ void __launch_bounds__(GROUP_THREADS)
coarse_upper_kernel(int64_t nc, const u32 *start, const u32 *rp, const u32 *adj, const u32 *grp, u32 *tmp_nb,
                    u32 *tmp_mult, u32 *up_cnt, u32 *low_cnt, int *overflow) {
  extern __shared__ u32 sm[];
  u32 *sk = sm, *heads = sm + GROUP_CAP; // heads[GROUP_CAP + 1]
  __shared__ u32 cnt_s, base_s, wsum[GROUP_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t cv = blockIdx.x; cv < nc; cv += gridDim.x) {
    const u32 gs = start[cv], ge = start[cv + 1];
    const u32 e0 = rp[gs], e1 = rp[ge];
    if (e1 - e0 > (u32)GROUP_CAP) {
      if (threadIdx.x == 0) *overflow = 1;
      continue;
    }
    if (threadIdx.x == 0) {
      cnt_s = 0;
      base_s = 0;
    }
    __syncthreads();
    for (u32 j = e0 + threadIdx.x; j < e1; j += blockDim.x) {
      const u32 t = adj[j];
      if (t >= gs) sk[atomicAdd(&cnt_s, 1u)] = grp[t]; // any order: sorted next
    }
    __syncthreads();
    const int cnt = (int)cnt_s;
    const int n2 = pow2_ceil(cnt > 1 ? cnt : 1);
    for (int i = cnt + threadIdx.x; i < n2; i += blockDim.x) sk[i] = 0xffffffffu;
    __syncthreads();
    block_bitonic_keys(sk, n2);
    // run heads -> heads[rank] = position
    for (int t0 = 0; t0 < cnt; t0 += blockDim.x) {
      const int i = t0 + threadIdx.x;
      const bool head = i < cnt && (i == 0 || sk[i] != sk[i - 1]);
      const unsigned bal = __ballot_sync(0xffffffffu, head);
      if (lane == 0) wsum[w] = __popc(bal);
      __syncthreads();
      u32 before = base_s;
      for (int q = 0; q < w; ++q) before += wsum[q];
      if (head) heads[before + __popc(bal & ((1u << lane) - 1))] = (u32)i;
      __syncthreads();
      if (threadIdx.x == 0) {
        u32 tot = 0;
        for (int q = 0; q < GROUP_THREADS / 32; ++q) tot += wsum[q];
        base_s += tot;
      }
      __syncthreads();
    }
    const int distinct = (int)base_s;
    if (threadIdx.x == 0) {
      heads[distinct] = (u32)cnt;
      up_cnt[cv] = (u32)distinct;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < distinct; d += blockDim.x) {
      const u32 key = sk[heads[d]];
      tmp_nb[e0 + d] = key;
      tmp_mult[e0 + d] = heads[d + 1] - heads[d];
      if (key != (u32)cv) atomicAdd(&low_cnt[key], 1u);
    }
    __syncthreads();
  }
}

__global__ void add_counts_kernel(int64_t nc, const u32 *a, const u32 *b, u32 *out, int *empty_vertex) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nc; v += (int64_t)gridDim.x * blockDim.x) {
    const u32 t = a[v] + b[v];
    out[v] = t;
    if (t == 0) *empty_vertex = 1; // the reference reads out of bounds here (Q8)
  }
}

// upper lists behind the lower part; mirrored entries into the targets' lower parts (unordered)
__global__ void coarse_fill_kernel(int64_t nc, const u32 *start, const u32 *rp, const u32 *tmp_nb, const u32 *tmp_mult,
                                   const u32 *up_cnt, const u32 *low_cnt, const u32 *crp, u32 *cursor, u32 *cadj,
                                   u32 *cmult) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t v = warp; v < nc; v += n_warps) {
    const u32 e0 = rp[start[v]], cnt = up_cnt[v], base = crp[v] + low_cnt[v];
    for (u32 d = lane; d < cnt; d += 32) {
      const u32 nb = tmp_nb[e0 + d], mu = tmp_mult[e0 + d];
      cadj[base + d] = nb;
      cmult[base + d] = mu;
      if (nb != (u32)v) {
        const u32 pos = crp[nb] + atomicAdd(&cursor[nb], 1u);
        cadj[pos] = (u32)v;
        cmult[pos] = mu;
      }
    }
  }
}

// sorts the first len[v] pairs of segment v (at off[v]) by key; one warp per segment
__global__ void __launch_bounds__(SORT_WARPS * 32)
sort_segments_kernel(int64_t n_seg, const u32 *off, const u32 *len, u32 *keys, u32 *vals, int *overflow) {
  extern __shared__ u32 sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  u32 *sk = sm + (size_t)w * 2 * WARP_CAP, *sv = sk + WARP_CAP;
  const int64_t warp = (int64_t)blockIdx.x * SORT_WARPS + w;
  const int64_t n_warps = (int64_t)gridDim.x * SORT_WARPS;
  for (int64_t v = warp; v < n_seg; v += n_warps) {
    const u32 o = off[v], l = len[v];
    if (l <= 1) continue;
    if (l <= 32) {
      u32 k = lane < l ? keys[o + lane] : 0xffffffffu, p = lane < l ? vals[o + lane] : 0;
      const u32 prev = __shfl_up_sync(0xffffffffu, k, 1);
      if (__any_sync(0xffffffffu, lane > 0 && lane < l && k < prev)) {
        sort32_pairs(k, p, lane);
        if (lane < l) {
          keys[o + lane] = k;
          vals[o + lane] = p;
        }
      }
    } else if (l <= (u32)WARP_CAP) {
      const int n2 = pow2_ceil((int)l);
      for (int i = lane; i < n2; i += 32) {
        sk[i] = i < (int)l ? keys[o + i] : 0xffffffffu;
        sv[i] = i < (int)l ? vals[o + i] : 0;
      }
      __syncwarp();
      warp_bitonic_pairs(sk, sv, n2, lane);
      for (int i = lane; i < (int)l; i += 32) {
        keys[o + i] = sk[i];
        vals[o + i] = sv[i];
      }
      __syncwarp();
    } else {
      if (lane == 0) *overflow = 1;
    }
  }
}

// ---- A' = P A P^T --------------------------------------------------------------------------
__global__ void inverse_perm_kernel(int64_t n, const u32 *perm, u32 *fwd, u32 *len2, const u32 *rp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const u32 o = perm[i];
    fwd[o] = (u32)i;
    len2[i] = rp[o + 1] - rp[o];
  }
}

// one warp per new row: columns mapped through the inverse permutation, sorted, values follow
template <typename T>
__global__ void __launch_bounds__(SORT_WARPS * 32)
reorder_rows_kernel(int64_t n, const u32 *perm, const u32 *fwd, const u32 *rp, const u32 *ci, const T *val,
                    const u32 *rp2, u32 *ci2, T *val2, int *overflow) {
  extern __shared__ u32 sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  u32 *sk = sm + (size_t)w * 2 * WARP_CAP, *sv = sk + WARP_CAP;
  const int64_t warp = (int64_t)blockIdx.x * SORT_WARPS + w;
  const int64_t n_warps = (int64_t)gridDim.x * SORT_WARPS;
  for (int64_t i = warp; i < n; i += n_warps) {
    const u32 o = perm[i];
    const u32 s = rp[o], l = rp[o + 1] - s, dst = rp2[i];
    if (l <= 32) {
      u32 k = lane < l ? fwd[ci[s + lane]] : 0xffffffffu, p = s + lane;
      const u32 prev = __shfl_up_sync(0xffffffffu, k, 1);
      if (__any_sync(0xffffffffu, lane > 0 && lane < l && k < prev)) sort32_pairs(k, p, lane);
      if (lane < l) {
        ci2[dst + lane] = k;
        val2[dst + lane] = val[p];
      }
    } else if (l <= (u32)WARP_CAP) {
      const int n2 = pow2_ceil((int)l);
      for (int t = lane; t < n2; t += 32) {
        sk[t] = t < (int)l ? fwd[ci[s + t]] : 0xffffffffu;
        sv[t] = s + t;
      }
      __syncwarp();
      warp_bitonic_pairs(sk, sv, n2, lane);
      for (int t = lane; t < (int)l; t += 32) {
        ci2[dst + t] = sk[t];
        val2[dst + t] = val[sv[t]];
      }
      __syncwarp();
    } else {
      if (lane == 0) *overflow = 1;
    }
  }
}

struct DevBuf { // frees what the build allocated, whichever way it leaves
  std::vector<void *> ptrs;
  ~DevBuf() {
    for (void *p : ptrs) cudaFree(p);
  }
  template <typename T> int get(T **p, size_t count) {
    void *q = nullptr;
    if (cudaMalloc(&q, (count ? count : 1) * sizeof(T)) != cudaSuccess) {
      cudaGetLastError();
      return HSPMV_ENOMEM;
    }
    ptrs.push_back(q);
    *p = (T *)q;
    return HSPMV_OK;
  }
};

int grid_of(int64_t items, int per_block, int sms, int per_sm) {
  int64_t g = (items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)sms * per_sm;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

} // namespace

int coarse_level1_device(int64_t n, int64_t nnz, const u32 *d_rp, const u32 *d_ci, const uvec &start,
                         int sms, cudaStream_t st, CoarseGraph &out, int64_t *launches_out) {
  const int64_t nc = (int64_t)start.size() - 1;
  if (nc < 1 || start.back() != (u32)n) return BANDK_USE_HOST;
  int64_t launches = 0;
  DevBuf B;
  u32 *d_start, *d_grp, *tmp_nb, *tmp_mult, *up_cnt, *low_cnt, *tot, *crp, *cursor;
  int *flags; // [0] overflow, [1] empty coarse vertex
  int rc;
  if ((rc = B.get(&d_start, (size_t)nc + 1)) || (rc = B.get(&d_grp, (size_t)n)) || (rc = B.get(&tmp_nb, (size_t)nnz)) ||
      (rc = B.get(&tmp_mult, (size_t)nnz)) || (rc = B.get(&up_cnt, (size_t)nc)) || (rc = B.get(&low_cnt, (size_t)nc)) ||
      (rc = B.get(&tot, (size_t)nc)) || (rc = B.get(&crp, (size_t)nc + 1)) || (rc = B.get(&cursor, (size_t)nc)) ||
      (rc = B.get(&flags, 2)))
    return rc == HSPMV_ENOMEM ? BANDK_USE_HOST : rc;
  BK_CU(cudaMemcpyAsync(d_start, start.data(), (size_t)(nc + 1) * 4, cudaMemcpyHostToDevice, st));
  BK_CU(cudaMemsetAsync(low_cnt, 0, (size_t)nc * 4, st));
  BK_CU(cudaMemsetAsync(cursor, 0, (size_t)nc * 4, st));
  BK_CU(cudaMemsetAsync(flags, 0, 8, st));

  group_of_row_kernel<<<grid_of(n, 256, sms, 16), 256, 0, st>>>(n, nc, d_start, d_grp);
  ++launches;
  const size_t smem_group = (size_t)(2 * GROUP_CAP + 1) * 4;
  BK_CU(cudaFuncSetAttribute(coarse_upper_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_group));
  coarse_upper_kernel<<<grid_of(nc, 1, sms, 3), GROUP_THREADS, smem_group, st>>>(nc, d_start, d_rp, d_ci, d_grp, tmp_nb,
                                                                              tmp_mult, up_cnt, low_cnt, flags);
  ++launches;
  add_counts_kernel<<<grid_of(nc, 256, sms, 16), 256, 0, st>>>(nc, up_cnt, low_cnt, tot, flags + 1);
  ++launches;
  if ((rc = scan_exclusive(tot, nc, crp, st, launches))) return rc;
  int h_flags[2];
  u32 cnnz = 0;
  BK_CU(cudaMemcpyAsync(h_flags, flags, 8, cudaMemcpyDeviceToHost, st));
  BK_CU(cudaMemcpyAsync(&cnnz, crp + nc, 4, cudaMemcpyDeviceToHost, st));
  BK_CU(cudaStreamSynchronize(st));
  if (h_flags[0]) return BANDK_USE_HOST;
  if (h_flags[1]) return HSPMV_EUNSUPPORTED;

  u32 *cadj, *cmult;
  if ((rc = B.get(&cadj, (size_t)cnnz)) || (rc = B.get(&cmult, (size_t)cnnz))) return rc == HSPMV_ENOMEM ? BANDK_USE_HOST : rc;
  coarse_fill_kernel<<<grid_of(nc, 8, sms, 8), 256, 0, st>>>(nc, d_start, d_rp, tmp_nb, tmp_mult, up_cnt, low_cnt, crp,
                                                             cursor, cadj, cmult);
  ++launches;
  const size_t smem_sort = (size_t)SORT_WARPS * 2 * WARP_CAP * 4;
  BK_CU(cudaFuncSetAttribute(sort_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sort));
  sort_segments_kernel<<<grid_of(nc, SORT_WARPS, sms, 3), SORT_WARPS * 32, smem_sort, st>>>(nc, crp, low_cnt, cadj, cmult,
                                                                                          flags);
  ++launches;
  out.n = nc;
  out.rp.resize((size_t)nc + 1);
  out.adj.resize((size_t)cnnz);
  out.mult.resize((size_t)cnnz);
  first_touch(out.adj.data(), (size_t)cnnz * 4); // uvec: nothing has written these pages yet
  first_touch(out.mult.data(), (size_t)cnnz * 4);
  BK_CU(cudaMemcpyAsync(out.rp.data(), crp, (size_t)(nc + 1) * 4, cudaMemcpyDeviceToHost, st));
  BK_CU(cudaMemcpyAsync(out.adj.data(), cadj, (size_t)cnnz * 4, cudaMemcpyDeviceToHost, st));
  BK_CU(cudaMemcpyAsync(out.mult.data(), cmult, (size_t)cnnz * 4, cudaMemcpyDeviceToHost, st));
  BK_CU(cudaMemcpyAsync(h_flags, flags, 8, cudaMemcpyDeviceToHost, st));
  BK_CU(cudaStreamSynchronize(st));
  if (h_flags[0]) return BANDK_USE_HOST; // a hub vertex with more than WARP_CAP lower neighbours
  if (launches_out) *launches_out += launches;
  return HSPMV_OK;
}

template <typename T>
static int reorder_device_impl(int64_t n, int64_t nnz, const u32 *d_perm, const u32 *d_rp, const u32 *d_ci,
                               const T *d_val, u32 *d_rp2, u32 *d_ci2, T *d_val2, int sms, cudaStream_t st,
                               int64_t *launches_out) {
  (void)nnz;
  int64_t launches = 0;
  DevBuf B;
  u32 *fwd, *len2;
  int *flag;
  int rc;
  if ((rc = B.get(&fwd, (size_t)n)) || (rc = B.get(&len2, (size_t)n)) || (rc = B.get(&flag, 1)))
    return rc == HSPMV_ENOMEM ? BANDK_USE_HOST : rc;
  BK_CU(cudaMemsetAsync(flag, 0, 4, st));
  inverse_perm_kernel<<<grid_of(n, 256, sms, 16), 256, 0, st>>>(n, d_perm, fwd, len2, d_rp);
  ++launches;
  if ((rc = scan_exclusive(len2, n, d_rp2, st, launches))) return rc;
  const size_t smem_sort = (size_t)SORT_WARPS * 2 * WARP_CAP * 4;
  BK_CU(cudaFuncSetAttribute(reorder_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sort));
  reorder_rows_kernel<T><<<grid_of(n, SORT_WARPS, sms, 3), SORT_WARPS * 32, smem_sort, st>>>(n, d_perm, fwd, d_rp, d_ci, d_val,
                                                                                           d_rp2, d_ci2, d_val2, flag);
  ++launches;
  int h_flag = 0;
  BK_CU(cudaMemcpyAsync(&h_flag, flag, 4, cudaMemcpyDeviceToHost, st));
  BK_CU(cudaStreamSynchronize(st));
  if (h_flag) return BANDK_USE_HOST;
  if (launches_out) *launches_out += launches;
  return HSPMV_OK;
}

int reorder_device(int dtype, int64_t n, int64_t nnz, const u32 *d_perm, const u32 *d_rp, const u32 *d_ci,
                   const void *d_val, u32 *d_rp2, u32 *d_ci2, void *d_val2, int sms, cudaStream_t st,
                   int64_t *launches) {
  if (dtype == HSPMV_F64)
    return reorder_device_impl<double>(n, nnz, d_perm, d_rp, d_ci, (const double *)d_val, d_rp2, d_ci2, (double *)d_val2,
                                       sms, st, launches);
  return reorder_device_impl<float>(n, nnz, d_perm, d_rp, d_ci, (const float *)d_val, d_rp2, d_ci2, (float *)d_val2, sms,
                                    st, launches);
}

} // namespace hspmv
