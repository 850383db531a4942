// C ABI of libhspmv.so (declared in include/hspmv.h): matrix handles, upload, Band-k build,
// kernel selection and launch, device-event timing.  Host index work lives in bandk.cpp,
// partition.cpp, io.cpp, heuristics.cpp; kernels in kernels.cuh / merge.cuh / gen.cuh.
#include "gen.cuh"
#include "host_index.h"
#include "kernels.cuh"
#include "merge.cuh"
#include "csrk_variants.h"
#include "bandk_gpu.h"

#include <chrono>
#include "sts.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

using hspmv::u32;
using hspmv::uvec;
using hspmv::ivec;

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      cudaGetLastError(); /* do not leave it for the next cudaGetLastError() */           \
      return fail(HSPMV_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    }                                                                                     \
  } while (0)

constexpr int64_t PAD = 16; // tail padding (elements) of every device array

struct DeviceProps {
  int device = -1, sms = 0, cc = 0;
  int64_t l2 = 0, hbm = 0;
  size_t max_smem_optin = 0;
  int max_persist_l2 = 0, max_window = 0;
};

int query_props(DeviceProps &p) {
  int dev;
  CU(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev));
  p.device = dev;
  p.sms = prop.multiProcessorCount;
  p.cc = prop.major * 10 + prop.minor;
  p.l2 = prop.l2CacheSize;
  p.hbm = (int64_t)prop.totalGlobalMem;
  p.max_smem_optin = prop.sharedMemPerBlockOptin;
  p.max_persist_l2 = prop.persistingL2CacheMaxSize;
  p.max_window = prop.accessPolicyMaxWindowSize;
  return HSPMV_OK;
}

} // namespace

struct hspmv_matrix {
  int dtype = HSPMV_F64;
  int64_t m = 0, n = 0, nnz = 0;
  int k = 1;
  int64_t n_sr = 0, n_ssr = 0;
  DeviceProps dev;

  int *d_rp = nullptr, *d_ci = nullptr;
  void *d_val = nullptr, *d_x = nullptr, *d_y = nullptr, *d_tmp = nullptr;
  u32 *d_sr = nullptr, *d_ssr = nullptr, *d_sr_e = nullptr, *d_perm = nullptr;
  // wall seconds of the last hspmv_build_csrk by phase (hspmv_build_phases): 0 host prep (row_ptr download, row
  // statistics, grouping), 1 level-1 coarse graph in HBM + its download, 2 RCM / level 2 / uncoarsening on the host
  // cores, 3 A' = P A P^T (permutation upload, reorder, row_ptr download), 4 maps + statistics of the new matrix, 5 total
  double build_phase[6] = {0, 0, 0, 0, 0, 0};
  u32 *d_dims = nullptr; // hspmv_device_view_get: {n_rows, n_cols, nnz, n_coarsest} as device words

  // CSR_VECTOR bins
  bool bins_ready = false;
  int *d_bin_rows = nullptr;
  int *d_long_seg = nullptr, *d_long_ptr = nullptr; // last bin: {slot, e0, e1} per segment, segments per row
  void *d_long_partial = nullptr;
  int64_t n_long_seg = 0;
  int64_t bin_ptr[hspmv::NUM_BINS + 1] = {0};
  int64_t stats[10] = {0};
  bool stats_ready = false;

  int64_t sr_max_rows = 0, sr_max_nnz = 0; // largest super-row (exact stage capacity)
  int64_t len_p90 = -1;                     // 90th percentile of the row lengths (ragged matrices only; lazily)

  // STS (hspmv_build_sts): level-set packs of the reordered matrix, diagonal positions, right-hand side, solution
  bool sts_ready = false;
  int64_t n_packs = 0, solo_packs = 0;
  int *d_diag = nullptr, *d_pack = nullptr, *d_ready = nullptr; // d_ready: n epoch words + 1 timeout word (barrier-free solve)
  int sts_epoch = 0;
  void *d_sts_b = nullptr, *d_sts_x = nullptr;
  std::vector<int32_t> h_pack;

  // hspmv_relabel_hub_columns: d_perm permutes x only (columns were renumbered, rows were not)
  bool cols_only_perm = false;
  // CSR_VECTOR bins run concurrently on auxiliary streams (independent rows): fork / join events
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};

  // hspmv_spmv_graph: `iters` products captured once as a CUDA graph
  cudaGraphExec_t graph_exec = nullptr;
  int graph_algo = -1, graph_iters = 0;
  int64_t graph_launches = 0;

  // CSR_BLOCK
  int64_t block_cap = 0, block_max_rows = 0, n_blocks = 0;
  int *d_blk_row = nullptr;

  // CSR_MERGE
  int64_t merge_tile = 0, merge_tiles = 0;
  int merge_ctas = 0;
  int *d_tile_row = nullptr, *d_carry_row = nullptr;
  void *d_carry_val = nullptr;

  // host copies (kept on request; needed for Band-k)
  ivec h_rp, h_ci;
  std::vector<unsigned char> h_val;
  uvec h_perm, h_sr, h_ssr;

  // pipelined host call (hspmv_spmv_host on large regular matrices): x pieces go up, row chunks
  // are computed as soon as the pieces they reference have landed, y chunks go down meanwhile
  struct Pipe {
    bool tried = false, ok = false;
    int chunks = 0;
    std::vector<int64_t> row_lo, x_lo; // chunk row bounds, x piece bounds (chunks+1 each)
    std::vector<int> need;             // last x piece chunk c reads
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> ev_x, ev_k;
  } pipe;

  int64_t opt[32] = {0};
  cudaStream_t stream = nullptr; // handle's own stream (timed runs, set_x/get_y)
  cudaStream_t last_stream = nullptr; // stream of the most recent hspmv_spmv (may be the caller's)
  bool last_stream_valid = false;
  void *d_flush = nullptr, *d_flush_rd = nullptr;
  size_t flush_bytes = 0;
  bool l2_window_set = false;
  std::string last_kernel;

  size_t vsize() const { return dtype == HSPMV_F64 ? 8 : 4; }
};

namespace {

void free_pipe(hspmv_matrix *A) {
  for (cudaEvent_t e : A->pipe.ev_x) cudaEventDestroy(e);
  for (cudaEvent_t e : A->pipe.ev_k) cudaEventDestroy(e);
  if (A->pipe.s_h2d) cudaStreamDestroy(A->pipe.s_h2d);
  if (A->pipe.s_d2h) cudaStreamDestroy(A->pipe.s_d2h);
  A->pipe = hspmv_matrix::Pipe();
}

void free_device(hspmv_matrix *A) {
  free_pipe(A);
  void *ptrs[] = {A->d_rp, A->d_ci, A->d_val, A->d_x, A->d_y, A->d_tmp, A->d_sr, A->d_ssr,
                  A->d_sr_e, A->d_perm, A->d_bin_rows, A->d_long_seg, A->d_long_ptr, A->d_long_partial,
                  A->d_tile_row, A->d_carry_row, A->d_blk_row, A->d_diag, A->d_pack, A->d_ready, A->d_sts_b, A->d_sts_x,
                  A->d_carry_val, A->d_flush, A->d_flush_rd, A->d_dims};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  for (int i = 0; i < 3; ++i) {
    if (A->aux[i]) cudaStreamDestroy(A->aux[i]);
    if (A->ev_join[i]) cudaEventDestroy(A->ev_join[i]);
  }
  if (A->ev_fork) cudaEventDestroy(A->ev_fork);
  if (A->graph_exec) cudaGraphExecDestroy(A->graph_exec);
  A->graph_exec = nullptr;
}

int alloc_vectors(hspmv_matrix *A) {
  const size_t vs = A->vsize();
  const int64_t big = (A->n > A->m ? A->n : A->m) + PAD;
  CU(cudaMalloc(&A->d_x, (size_t)(A->n + PAD) * vs));
  CU(cudaMalloc(&A->d_y, (size_t)(A->m + PAD) * vs));
  CU(cudaMalloc(&A->d_tmp, (size_t)big * vs));
  CU(cudaMemset(A->d_x, 0, (size_t)(A->n + PAD) * vs));
  CU(cudaMemset(A->d_y, 0, (size_t)(A->m + PAD) * vs));
  return HSPMV_OK;
}

int alloc_matrix_arrays(hspmv_matrix *A) {
  const size_t vs = A->vsize();
  CU(cudaMalloc(&A->d_rp, (size_t)(A->m + 1 + PAD) * 4));
  CU(cudaMalloc(&A->d_ci, (size_t)(A->nnz + PAD) * 4));
  CU(cudaMalloc(&A->d_val, (size_t)(A->nnz + PAD) * vs));
  // the pads are read by aligned quads / TMA windows that straddle the end: keep them defined
  CU(cudaMemset((char *)A->d_rp + (size_t)(A->m + 1) * 4, 0, PAD * 4));
  CU(cudaMemset((char *)A->d_ci + (size_t)A->nnz * 4, 0, PAD * 4));
  CU(cudaMemset((char *)A->d_val + (size_t)A->nnz * vs, 0, PAD * vs));
  return HSPMV_OK;
}

int new_handle(hspmv_matrix **out, int dtype, int64_t m, int64_t n, int64_t nnz) {
  if (!out) return fail(HSPMV_EINVAL, "null handle pointer");
  if (dtype != HSPMV_F32 && dtype != HSPMV_F64) return fail(HSPMV_EINVAL, "dtype must be HSPMV_F32 or HSPMV_F64");
  if (m < 0 || n < 0 || nnz < 0 || nnz > 2147483647LL - PAD || m > 2147483647LL - PAD)
    return fail(HSPMV_EINVAL, "sizes out of the 32-bit index range of the file formats");
  hspmv_matrix *A = new hspmv_matrix();
  A->dtype = dtype;
  A->m = m;
  A->n = n;
  A->nnz = nnz;
  int rc = query_props(A->dev);
  if (rc) {
    delete A;
    return rc;
  }
  if (A->dev.cc < 100) {
    delete A;
    return fail(HSPMV_ECUDA, "libhspmv needs an sm_100 (Blackwell) device; there is no fallback path");
  }
  cudaError_t e = cudaStreamCreateWithFlags(&A->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete A;
    return fail(HSPMV_ECUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
  }
  *out = A;
  return HSPMV_OK;
}

int grid_for(int64_t items, int threads, int cap) {
  int64_t g = (items + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

#define LAUNCHED() g_launches.fetch_add(1, std::memory_order_relaxed)

// ---- bins ---------------------------------------------------------------------------
// host row pointers of the handle: the kept copy itself (no 256 MB duplicate at config 5), else a download into `scratch`
int ensure_host_rp(hspmv_matrix *A, std::vector<int32_t> &scratch, const int32_t *&rp) {
  if (!A->h_rp.empty()) {
    rp = A->h_rp.data();
    return HSPMV_OK;
  }
  scratch.resize((size_t)A->m + 1);
  CU(cudaMemcpy(scratch.data(), A->d_rp, (size_t)(A->m + 1) * 4, cudaMemcpyDeviceToHost));
  rp = scratch.data();
  return HSPMV_OK;
}

int ensure_stats(hspmv_matrix *A) {
  if (A->stats_ready) return HSPMV_OK;
  std::vector<int32_t> scratch;
  const int32_t *rp = nullptr;
  int rc = ensure_host_rp(A, scratch, rp);
  if (rc) return rc;
  hspmv::row_stats(A->m, rp, A->stats);
  A->stats_ready = true;
  return HSPMV_OK;
}

int ensure_bins(hspmv_matrix *A) {
  if (A->bins_ready) return HSPMV_OK;
  std::vector<int32_t> scratch, rows;
  const int32_t *rp = nullptr;
  int rc = ensure_host_rp(A, scratch, rp);
  if (rc) return rc;
  hspmv::bin_rows(A->m, rp, rows, A->bin_ptr);
  if (A->d_bin_rows) cudaFree(A->d_bin_rows);
  CU(cudaMalloc(&A->d_bin_rows, (size_t)(A->m + 1) * 4));
  if (A->m) CU(cudaMemcpy(A->d_bin_rows, rows.data(), (size_t)A->m * 4, cudaMemcpyHostToDevice));
  // rows of the last bin are cut into segments of LONG_SEG entries, one CTA each
  const int64_t LONG_SEG = 16384;
  std::vector<int32_t> seg, seg_ptr(1, 0);
  for (int64_t i = A->bin_ptr[hspmv::NUM_BINS - 1]; i < A->bin_ptr[hspmv::NUM_BINS]; ++i) {
    const int32_t r = rows[(size_t)i];
    const int64_t e0 = rp[(size_t)r], e1 = rp[(size_t)r + 1];
    const int64_t pieces = (e1 - e0 + LONG_SEG - 1) / LONG_SEG;
    for (int64_t p = 0; p < pieces; ++p) {
      seg.push_back((int32_t)(i - A->bin_ptr[hspmv::NUM_BINS - 1]));
      seg.push_back((int32_t)(e0 + (e1 - e0) * p / pieces));
      seg.push_back((int32_t)(e0 + (e1 - e0) * (p + 1) / pieces));
    }
    seg_ptr.push_back((int32_t)(seg.size() / 3));
  }
  for (void *q : {(void *)A->d_long_seg, (void *)A->d_long_ptr, A->d_long_partial})
    if (q) cudaFree(q);
  A->d_long_seg = A->d_long_ptr = nullptr;
  A->d_long_partial = nullptr;
  A->n_long_seg = (int64_t)seg.size() / 3;
  if (A->n_long_seg) {
    CU(cudaMalloc(&A->d_long_seg, seg.size() * 4));
    CU(cudaMalloc(&A->d_long_ptr, seg_ptr.size() * 4));
    CU(cudaMalloc(&A->d_long_partial, (size_t)A->n_long_seg * A->vsize()));
    CU(cudaMemcpy(A->d_long_seg, seg.data(), seg.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(A->d_long_ptr, seg_ptr.data(), seg_ptr.size() * 4, cudaMemcpyHostToDevice));
  }
  A->bins_ready = true;
  return HSPMV_OK;
}

// ---- launches -------------------------------------------------------------------------
template <typename T, int VEC>
void launch_vector(hspmv_matrix *A, bool list, const int *rows, int64_t n_rows, int64_t row_base,
                   const T *x, T *y, cudaStream_t st) {
  if (n_rows <= 0) return;
  const int threads = 256;
  const int64_t groups_per_cta = threads / VEC;
  int64_t want = (n_rows + groups_per_cta - 1) / groups_per_cta;
  const int64_t cap = (int64_t)A->dev.sms * 64; // grid-stride beyond this
  int grid = (int)(want < cap ? want : cap);
  if (list)
    hspmv::csr_vector_kernel<T, VEC, true><<<grid, threads, 0, st>>>(
        n_rows, rows, 0, A->d_rp, A->d_ci, (const T *)A->d_val, x, y);
  else
    hspmv::csr_vector_kernel<T, VEC, false><<<grid, threads, 0, st>>>(
        n_rows, nullptr, row_base, A->d_rp, A->d_ci, (const T *)A->d_val, x, y);
  LAUNCHED();
}

template <typename T>
void launch_vector_vec(hspmv_matrix *A, int vec, bool list, const int *rows, int64_t n_rows,
                       int64_t row_base, const T *x, T *y, cudaStream_t st) {
  switch (vec) {
  case 1: launch_vector<T, 1>(A, list, rows, n_rows, row_base, x, y, st); break;
  case 2: launch_vector<T, 2>(A, list, rows, n_rows, row_base, x, y, st); break;
  case 4: launch_vector<T, 4>(A, list, rows, n_rows, row_base, x, y, st); break;
  case 8: launch_vector<T, 8>(A, list, rows, n_rows, row_base, x, y, st); break;
  case 16: launch_vector<T, 16>(A, list, rows, n_rows, row_base, x, y, st); break;
  default: launch_vector<T, 32>(A, list, rows, n_rows, row_base, x, y, st); break;
  }
}

// CSR_VECTOR: one contiguous launch when every row fits the chosen lane count in a few steps,
// otherwise one launch per non-empty bin.
template <typename T>
int run_csr_vector(hspmv_matrix *A, const T *x, T *y, cudaStream_t st, int64_t row_lo = 0, int64_t row_hi = -1) {
  int rc = ensure_stats(A);
  if (rc) return rc;
  const int64_t max_len = A->stats[8];
  const double mean = A->m ? (double)A->nnz / (double)A->m : 0.0;
  int vec = (int)A->opt[HSPMV_OPT_VEC];
  if (vec <= 0) { // lanes * 4 entries per step should cover the mean row (+3 for misalignment)
    vec = 1;
    while (vec < 32 && vec * 4 < mean + 3) vec <<= 1;
  }
  if (row_hi < 0) row_hi = A->m;
  const bool ranged = row_lo != 0 || row_hi != A->m;
  const bool regular = max_len <= (int64_t)vec * 4 * 8 || A->opt[HSPMV_OPT_VEC] > 0 || ranged;
  char name[96];
  if (regular) {
    launch_vector_vec<T>(A, vec, false, nullptr, row_hi - row_lo, row_lo, x, y, st);
    std::snprintf(name, sizeof(name), "csr_vector<%s,vec=%d> contiguous", sizeof(T) == 8 ? "f64" : "f32", vec);
    A->last_kernel = name;
    return HSPMV_OK;
  }
  rc = ensure_bins(A);
  if (rc) return rc;
  static const int bin_vec[hspmv::NUM_BINS] = {1, 1, 2, 4, 8, 16, 32, 32};
  // The bins cover disjoint rows, so their launches are independent: they go round-robin to the caller's stream and
  // three auxiliary streams (fork / join with events), and the tail of one bin's grid overlaps the next bin's start
  // instead of idling the GPU eight times per product.
  if (!A->ev_fork) {
    CU(cudaEventCreateWithFlags(&A->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < 3; ++i) {
      CU(cudaStreamCreateWithFlags(&A->aux[i], cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&A->ev_join[i], cudaEventDisableTiming));
    }
  }
  CU(cudaEventRecord(A->ev_fork, st));
  for (int i = 0; i < 3; ++i) CU(cudaStreamWaitEvent(A->aux[i], A->ev_fork, 0));
  int launches = 0, lane = 0;
  // largest bins first (they take longest), the long-row pair stays on one stream
  int order[hspmv::NUM_BINS];
  for (int b = 0; b < hspmv::NUM_BINS; ++b) order[b] = b;
  std::sort(order, order + hspmv::NUM_BINS, [&](int a, int b) {
    const int64_t wa = (A->bin_ptr[a + 1] - A->bin_ptr[a]) * (int64_t)bin_vec[a] * (a == hspmv::NUM_BINS - 1 ? 64 : 1);
    const int64_t wb = (A->bin_ptr[b + 1] - A->bin_ptr[b]) * (int64_t)bin_vec[b] * (b == hspmv::NUM_BINS - 1 ? 64 : 1);
    return wa > wb;
  });
  for (int oi = 0; oi < hspmv::NUM_BINS; ++oi) {
    const int b = order[oi];
    const int64_t cnt = A->bin_ptr[b + 1] - A->bin_ptr[b];
    if (!cnt) continue;
    cudaStream_t s_ = lane == 0 ? st : A->aux[lane - 1];
    lane = (lane + 1) & 3;
    const int *rows = A->d_bin_rows + A->bin_ptr[b];
    if (b == 0) {
      hspmv::zero_rows_kernel<T><<<grid_for(cnt, 256, A->dev.sms * 8), 256, 0, s_>>>(cnt, rows, y);
      LAUNCHED();
    } else if (b == hspmv::NUM_BINS - 1) {
      const int64_t ns = A->n_long_seg;
      hspmv::csr_longrow_kernel<T><<<(int)(ns < A->dev.sms * 8 ? ns : A->dev.sms * 8), 256, 0, s_>>>(
          ns, A->d_long_seg, A->d_ci, (const T *)A->d_val, x, (T *)A->d_long_partial);
      LAUNCHED();
      hspmv::longrow_reduce_kernel<T><<<grid_for(cnt, 256, A->dev.sms * 8), 256, 0, s_>>>(cnt, rows, A->d_long_ptr,
                                                                                         (const T *)A->d_long_partial, y);
      LAUNCHED();
    } else {
      launch_vector_vec<T>(A, bin_vec[b], true, rows, cnt, 0, x, y, s_);
    }
    ++launches;
  }
  for (int i = 0; i < 3; ++i) {
    CU(cudaEventRecord(A->ev_join[i], A->aux[i]));
    CU(cudaStreamWaitEvent(st, A->ev_join[i], 0));
  }
  std::snprintf(name, sizeof(name), "csr_vector<%s> binned (%d launches on 4 streams)", sizeof(T) == 8 ? "f64" : "f32", launches);
  A->last_kernel = name;
  return HSPMV_OK;
}

// largest compiled MAIN <= want_main for this vec (want_main <= 0: the fully masked loop)
hspmv::CsrkVariant pick_variant(int vec, int want_main, int want_tail = 0) {
  hspmv::CsrkVariant best = {vec, 0, 4};
  bool have = false;
  for (int i = 0; i < hspmv::kNumCsrkVariants; ++i) {
    const hspmv::CsrkVariant &c = hspmv::kCsrkVariants[i];
    if (c.vec != vec) continue;
    if (have && c.main == best.main) continue; // the first entry of a (vec, main) pair is its default
    if (!have || (c.main <= want_main && c.main > best.main) || (best.main > want_main && c.main < best.main)) {
      if (!have || c.main <= want_main || best.main > want_main) best = c;
      have = true;
    }
  }
  if (want_tail > 0) { // smallest compiled tail >= the request for that (vec, main), else the largest
    hspmv::CsrkVariant alt = best;
    bool found = false;
    for (int i = 0; i < hspmv::kNumCsrkVariants; ++i) {
      const hspmv::CsrkVariant &c = hspmv::kCsrkVariants[i];
      if (c.vec != vec || c.main != best.main) continue;
      if (c.tail >= want_tail ? (!found || alt.tail < want_tail || c.tail < alt.tail) : (!found || (alt.tail < want_tail && c.tail > alt.tail))) {
        alt = c;
        found = true;
      }
    }
    best = alt;
  }
  return best;
}

template <typename T>
int launch_csrk(hspmv_matrix *A, const hspmv::CsrkParams<T> &p, int grid, int warps, size_t smem,
                cudaStream_t st, const hspmv::CsrkVariant &v) {
  const int rc = hspmv::launch_csrk_variant(p, grid, warps, smem, st, A->dev.device, v);
  if (rc == -1) return fail(HSPMV_EINVAL, "no compiled csrk_tma variant for this (vec, chunk)");
  if (rc) return fail(HSPMV_ECUDA, std::string("csrk_tma launch: ") + cudaGetErrorString((cudaError_t)rc));
  LAUNCHED();
  return HSPMV_OK;
}

// CSRK (k = 2,3 after hspmv_build_csrk / hspmv_create_csrk) and CSR_STREAM (k = 1, fixed-row tiles)
struct HaloWait { // hspmv_spmv_halo: three row ranges + flags the boundary tiles wait for
  int n_ranges = 0;
  int64_t lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  const int *flags = nullptr;
  int n_wait = 0, step = 0;
  int *timed_out = nullptr;
};

template <typename T>
int run_csrk(hspmv_matrix *A, bool stream_mode, const T *x, T *y, cudaStream_t st, int64_t row_lo = 0, int64_t row_hi = -1,
             const HaloWait *halo = nullptr) {
  if (!stream_mode && A->k < 2) return fail(HSPMV_ESTATE, "HSPMV_ALGO_CSRK needs hspmv_build_csrk or hspmv_create_csrk first");
  int rc0 = ensure_stats(A);
  if (rc0) return rc0;
  const double mean = A->m ? (double)A->nnz / (double)A->m : 1.0;
  const int64_t max_len = A->stats[8];
  // lanes per row: one round of gathers per lane should cover the typical row
  int vec = (int)A->opt[HSPMV_OPT_VEC];
  int chunk = (int)A->opt[HSPMV_OPT_CHUNK];
  int tail = (int)A->opt[HSPMV_OPT_TAIL];
  // Ragged rows (config 3: 5-dof shell blocks, 35 +- 10 per row): an unmasked chunk sized for the
  // mean row sends every shorter or longer row through extra rounds of gather latency.  One
  // masked round of `tail` gathers per lane that covers the 90th-percentile row is 40 % faster
  // (profiles/r1_sweep_ragged_cfg3.jsonl), with one row per lane group per tile so that the
  // register-resident single-pass path applies.
  const bool ragged = vec <= 0 && chunk == 0 && tail == 0 && (double)max_len > mean * 1.15 + 1.0;
  int ragged_rpt = 0;
  if (ragged) {
    if (A->len_p90 < 0) {
      std::vector<int32_t> scratch;
      const int32_t *rp = nullptr;
      int rc = ensure_host_rp(A, scratch, rp);
      if (rc) return rc;
      A->len_p90 = hspmv::row_len_percentile(A->m, rp, 90);
    }
    const int64_t cover = A->len_p90 > 4 ? A->len_p90 : 4;
    vec = 1;
    while (vec < 32 && (double)cover / vec > 12.0) vec <<= 1;
    tail = (int)((cover + vec - 1) / vec);
    if (tail < 4) tail = 4;
    chunk = -1;
    ragged_rpt = 32 / vec;
    while (ragged_rpt * mean < 200.0) ragged_rpt += 32 / vec; // keep a TMA tile above ~2.4 KB
  }
  if (vec <= 0) { // smallest group whose lanes each get <= 14 nonzeros of a typical row (one unmasked chunk)
    vec = 1;
    while (vec < 32 && mean / vec > 14.0) vec <<= 1;
  }
  // gathers a lane keeps in flight per unmasked chunk: the whole share of a typical row if it is
  // short enough (then a row costs one round of latency), else a divisor-ish of it.
  // HSPMV_OPT_CHUNK: > 0 requests that chunk, < 0 forces the fully masked loop.
  if (chunk == 0) {
    const int per_lane = (int)(mean / vec);           // floor: every lane of the group has that many
    chunk = per_lane <= 13 ? per_lane : (per_lane % 9 == 0 ? 9 : 8);
    if (chunk < 3) chunk = -1;
  }
  const hspmv::CsrkVariant variant = pick_variant(vec, chunk, tail);
  int warps = (int)A->opt[HSPMV_OPT_WARPS];
  // r1 sweep: occupancy matters more than stage size; r2: the shapes that fit 96 registers run 18 warps (640-thread build)
  // ... but only for a launch over the whole matrix: on the 8-GPU shards (hspmv_spmv_halo, 8 M rows per rank) 18 warps
  // measured 6 % SLOWER than 16 on one box in one run (0.4463 vs 0.4201 ms per step, profiles/r2_ab_n8_warps_xupdate.txt),
  // so halo launches and row-range launches keep the 16-warp form that every multi-GPU number of this round was taken with
  const bool whole = !halo && (!stream_mode || (row_lo == 0 && (row_hi < 0 || row_hi == A->m)));
  const bool wide_ok = whole && hspmv::csrk_has_wide(variant);
  if (warps <= 0) warps = wide_ok ? hspmv::CSRK_WIDE_WARPS : 16;
  if (warps > (wide_ok ? 20 : 16)) warps = wide_ok ? 20 : 16;
  int stages = (int)A->opt[HSPMV_OPT_STAGES];
  if (stages <= 0) stages = 2;

  if (row_hi < 0) row_hi = A->m;
  hspmv::CsrkParams<T> p;
  p.m = A->m;
  p.n_ranges = 1;
  p.range_lo[0] = stream_mode ? row_lo : 0;
  p.range_hi[0] = stream_mode ? row_hi : A->m;
  p.range_lo[1] = p.range_hi[1] = p.range_lo[2] = p.range_hi[2] = 0;
  p.wait_flags = nullptr;
  p.n_wait = 0;
  p.wait_step = 0;
  p.timed_out = nullptr;
  if (stream_mode && halo) {
    p.n_ranges = halo->n_ranges;
    for (int i = 0; i < halo->n_ranges; ++i) {
      p.range_lo[i] = halo->lo[i];
      p.range_hi[i] = halo->hi[i];
    }
    p.wait_flags = halo->flags;
    p.n_wait = halo->n_wait;
    p.wait_step = halo->step;
    p.timed_out = halo->timed_out;
  }
  p.rp = A->d_rp;
  p.ci = A->d_ci;
  p.val = (const T *)A->d_val;
  p.x = x;
  p.y = y;
  p.ssr_ptr = A->d_ssr;
  p.sr_ptr = A->d_sr;
  p.sr_e = A->d_sr_e;
  p.n_ssr = A->n_ssr;
  p.stages = stages;
  p.rows_per_tile = 0;

  int64_t typical_rows, typical_nnz, worst_nnz;
  if (stream_mode) {
    int rpt = (int)A->opt[HSPMV_OPT_ROWS_PER_TILE];
    if (rpt <= 0 && ragged_rpt > 0) rpt = ragged_rpt;
    if (rpt <= 0) { // whole passes of the warp, ~432 nonzeros per tile (r1 sweep: 16 warps x 2 stages of that)
      const int per_pass = 32 / vec;
      int passes = (int)(432.0 / ((mean > 1 ? mean : 1.0) * per_pass) + 0.5);
      if (passes < 1) passes = 1;
      rpt = passes * per_pass;
    }
    p.k = 1;
    p.rows_per_tile = rpt;
    p.tiles_before[0] = 0;
    for (int i = 0; i < 3; ++i) {
      const int64_t len = i < p.n_ranges ? p.range_hi[i] - p.range_lo[i] : 0;
      p.tiles_before[i + 1] = p.tiles_before[i] + (len + rpt - 1) / rpt;
    }
    p.n_sr = p.tiles_before[3];
    typical_rows = rpt;
    typical_nnz = (int64_t)(rpt * mean + 0.999);
    worst_nnz = (int64_t)rpt * max_len; // a tile can never hold more than this
  } else {
    p.k = A->k;
    p.n_sr = A->n_sr;
    typical_rows = A->n_sr ? (A->m + A->n_sr - 1) / A->n_sr : 1;
    typical_nnz = A->n_sr ? (A->nnz + A->n_sr - 1) / A->n_sr : 1;
    worst_nnz = A->sr_max_nnz > 0 ? A->sr_max_nnz : typical_nnz * 4;
  }

  // Stage capacity: explicit option, else 1.5x the typical super-row (never more than the worst
  // case).  If warps*stages of that do not fit in shared memory, give up warps first, capacity
  // last: an under-sized stage sends every super-row down the unstaged path.
  int cap_nnz = (int)A->opt[HSPMV_OPT_STAGE_NNZ];
  if (cap_nnz <= 0) {
    int64_t c = typical_nnz * 3 / 2 + 64;
    if (c > worst_nnz) c = worst_nnz;
    if (c < typical_nnz) c = typical_nnz;
    cap_nnz = (int)((c + 3) & ~(int64_t)3);
  }
  if (cap_nnz < 64) cap_nnz = 64;
  cap_nnz = (cap_nnz + 3) & ~3;
  int64_t want_rows = stream_mode ? typical_rows : typical_rows * 2 + 32;
  if (!stream_mode && A->sr_max_rows > 0 && A->sr_max_rows < want_rows) want_rows = A->sr_max_rows;
  int cap_rows = (int)((want_rows + 3) & ~(int64_t)3);
  if (cap_rows < 8) cap_rows = 8;
  const size_t budget = A->dev.max_smem_optin - 1024;
  auto total = [&](int w, int cn, int cr) {
    return (((size_t)w * stages * 8 + 127) & ~(size_t)127) +
           (size_t)w * stages * hspmv::csrk_stage_bytes(cn, cr, sizeof(T));
  };
  // 18 warps only while their stages leave the x gathers ~30 KB of L1: with 19 warps (205 KB of stages) or with CSR-2's
  // larger stages at 18 (202 KB) the kernel is 6-20 % SLOWER than with 16 (profiles/r2_ab_warps_per_sm.txt)
  if (warps > 16 && !A->opt[HSPMV_OPT_WARPS] && total(warps, cap_nnz, cap_rows) > (size_t)198000) warps = 16;
  while (warps > 4 && total(warps, cap_nnz, cap_rows) > budget) --warps;
  while (total(warps, cap_nnz, cap_rows) > budget && cap_nnz > 64) cap_nnz = (cap_nnz * 3 / 4) & ~3;
  while (total(warps, cap_nnz, cap_rows) > budget && cap_rows > 8) cap_rows = (cap_rows / 2) & ~3;
  while (warps > 1 && total(warps, cap_nnz, cap_rows) > budget) --warps;
  if (total(warps, cap_nnz, cap_rows) > budget) return fail(HSPMV_EINVAL, "stages do not fit shared memory");
  p.cap_nnz = cap_nnz;
  p.cap_rows = cap_rows;
  const size_t smem = total(warps, cap_nnz, cap_rows);

  const int64_t work_items = (!stream_mode && A->k == 3) ? A->n_ssr : (p.n_sr + warps - 1) / warps;
  int ctas_per_sm = (int)A->opt[HSPMV_OPT_CTAS_PER_SM];
  if (ctas_per_sm <= 0) {
    ctas_per_sm = (int)((A->dev.max_smem_optin + 1024) / (smem + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm * warps > 20) ctas_per_sm = 20 / warps > 0 ? 20 / warps : 1; // register file: <= 20 warps of <= 100 regs
  }
  int64_t grid = (int64_t)A->dev.sms * ctas_per_sm;
  if (grid > work_items) grid = work_items;
  if (grid < 1) grid = 1;

  char name[192];
  int rc = launch_csrk<T>(A, p, (int)grid, warps, smem, st, variant);
  if (rc) return rc;
  std::snprintf(name, sizeof(name), "csrk_tma<%s,vec=%d,main=%d,tail=%d> k=%d grid=%d warps=%d stages=%d cap_nnz=%d cap_rows=%d%s",
                sizeof(T) == 8 ? "f64" : "f32", vec, variant.main, variant.tail, p.k, (int)grid, warps, stages, cap_nnz, cap_rows,
                stream_mode ? " (CSR stream tiles)" : "");
  A->last_kernel = name;
  return HSPMV_OK;
}

// CSR_BLOCK: one CTA per row block of <= cap nonzeros, products staged in shared memory (kernels.cuh)
template <typename T, int VEC>
int launch_csr_block(hspmv_matrix *A, int cap, size_t smem, const T *x, T *y, cudaStream_t st) {
  auto kern = hspmv::csr_block_kernel<T, VEC, 256>;
  static size_t configured[64] = {0};
  size_t &have = configured[A->dev.device & 63];
  if (smem > have) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    have = smem;
  }
  kern<<<(int)A->n_blocks, 256, smem, st>>>((int)A->n_blocks, cap, A->d_blk_row, A->d_rp, A->d_ci, (const T *)A->d_val, x, y);
  LAUNCHED();
  return HSPMV_OK;
}

template <typename T>
int run_csr_block(hspmv_matrix *A, const T *x, T *y, cudaStream_t st) {
  const double mean = A->m ? (double)A->nnz / (double)A->m : 0.0;
  int64_t cap = A->opt[HSPMV_OPT_BLOCK_NNZ];
  if (cap <= 0) cap = 2048;
  cap = (cap + 3) & ~(int64_t)3;
  int64_t max_rows = cap; // a block never needs more row slots than nonzero slots + empty rows; bounded for shared memory
  if (max_rows > 2048) max_rows = 2048;
  if (A->block_cap != cap || A->block_max_rows != max_rows || !A->d_blk_row) {
    std::vector<int32_t> scratch, blk;
    const int32_t *rp = nullptr;
    int rc = ensure_host_rp(A, scratch, rp);
    if (rc) return rc;
    hspmv::block_rows(A->m, rp, cap, max_rows, blk);
    if (A->d_blk_row) cudaFree(A->d_blk_row);
    A->d_blk_row = nullptr;
    CU(cudaMalloc(&A->d_blk_row, blk.size() * 4 + 16));
    CU(cudaMemcpy(A->d_blk_row, blk.data(), blk.size() * 4, cudaMemcpyHostToDevice));
    A->n_blocks = (int64_t)blk.size() - 1;
    A->block_cap = cap;
    A->block_max_rows = max_rows;
  }
  if (A->n_blocks <= 0) return HSPMV_OK;
  int vec = (int)A->opt[HSPMV_OPT_VEC];
  if (vec <= 0) { // lanes per row in the reduction phase: a lane adds <= 24 products of a typical row (r2 sweep: fewer lanes win)
    vec = 1;
    while (vec < 32 && mean / vec > 24.0) vec <<= 1;
  }
  const size_t smem = (size_t)(cap + 8) * sizeof(T) + (size_t)(max_rows + 1) * 4;
  int rc;
  switch (vec) {
  case 1: rc = launch_csr_block<T, 1>(A, (int)cap, smem, x, y, st); break;
  case 2: rc = launch_csr_block<T, 2>(A, (int)cap, smem, x, y, st); break;
  case 4: rc = launch_csr_block<T, 4>(A, (int)cap, smem, x, y, st); break;
  case 8: rc = launch_csr_block<T, 8>(A, (int)cap, smem, x, y, st); break;
  case 16: rc = launch_csr_block<T, 16>(A, (int)cap, smem, x, y, st); break;
  default: rc = launch_csr_block<T, 32>(A, (int)cap, smem, x, y, st); vec = 32; break;
  }
  if (rc) return rc;
  char name[128];
  std::snprintf(name, sizeof(name), "csr_block<%s,vec=%d> cap=%d blocks=%lld threads=256", sizeof(T) == 8 ? "f64" : "f32", vec, (int)cap,
                (long long)A->n_blocks);
  A->last_kernel = name;
  return HSPMV_OK;
}

template <typename T>
int run_csr_merge(hspmv_matrix *A, const T *x, T *y, cudaStream_t st) {
  int64_t tile = A->opt[HSPMV_OPT_MERGE_TILE];
  if (tile <= 0) tile = 4096;
  tile &= ~(int64_t)3;
  int ctas_per_sm = (int)A->opt[HSPMV_OPT_CTAS_PER_SM];
  if (ctas_per_sm <= 0) ctas_per_sm = 4;
  if (A->nnz == 0) {
    CU(cudaMemsetAsync(y, 0, (size_t)A->m * sizeof(T), st));
    A->last_kernel = "memset (empty matrix)";
    return HSPMV_OK;
  }
  const int64_t n_tiles = (A->nnz + tile - 1) / tile;
  int64_t ctas = (int64_t)A->dev.sms * ctas_per_sm;
  if (ctas > n_tiles) ctas = n_tiles;
  if (A->merge_tile != tile || A->merge_ctas != (int)ctas) {
    if (A->d_tile_row) cudaFree(A->d_tile_row);
    if (A->d_carry_row) cudaFree(A->d_carry_row);
    if (A->d_carry_val) cudaFree(A->d_carry_val);
    A->d_tile_row = A->d_carry_row = nullptr;
    A->d_carry_val = nullptr;
    CU(cudaMalloc(&A->d_tile_row, (size_t)(n_tiles + 1) * 4));
    CU(cudaMalloc(&A->d_carry_row, (size_t)ctas * 4));
    CU(cudaMalloc(&A->d_carry_val, (size_t)ctas * 8));
    hspmv::tile_rows_kernel<<<grid_for(n_tiles + 1, 256, A->dev.sms * 8), 256, 0, st>>>(
        A->m, A->d_rp, tile, n_tiles, A->d_tile_row);
    LAUNCHED();
    A->merge_tile = tile;
    A->merge_tiles = n_tiles;
    A->merge_ctas = (int)ctas;
  }
  const size_t smem = (size_t)tile * sizeof(T);
  auto kern = hspmv::csr_merge_kernel<T, 256>;
  static size_t configured[64] = {0};
  size_t &have = configured[A->dev.device & 63];
  if (smem > have) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    have = smem;
  }
  kern<<<(int)ctas, 256, smem, st>>>(A->m, A->nnz, (int)tile, n_tiles, A->d_tile_row, A->d_rp, A->d_ci,
                                     (const T *)A->d_val, x, y, A->d_carry_row, (T *)A->d_carry_val);
  LAUNCHED();
  hspmv::merge_fixup_kernel<T><<<1, 32, 0, st>>>((int)ctas, A->d_carry_row, (const T *)A->d_carry_val, y);
  LAUNCHED();
  char name[96];
  std::snprintf(name, sizeof(name), "csr_merge<%s> tile=%d ctas=%d + fixup", sizeof(T) == 8 ? "f64" : "f32",
                (int)tile, (int)ctas);
  A->last_kernel = name;
  return HSPMV_OK;
}

int apply_l2_window(hspmv_matrix *A, const void *x, cudaStream_t st) {
  if (!A->opt[HSPMV_OPT_L2_PERSIST_X] || A->dev.max_persist_l2 <= 0) return HSPMV_OK;
  size_t bytes = (size_t)A->n * A->vsize();
  size_t persist = bytes < (size_t)A->dev.max_persist_l2 ? bytes : (size_t)A->dev.max_persist_l2;
  if (!A->l2_window_set) {
    CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist));
    A->l2_window_set = true;
  }
  cudaStreamAttrValue attr;
  std::memset(&attr, 0, sizeof(attr));
  size_t win = bytes < (size_t)A->dev.max_window ? bytes : (size_t)A->dev.max_window;
  attr.accessPolicyWindow.base_ptr = const_cast<void *>(x);
  attr.accessPolicyWindow.num_bytes = win;
  attr.accessPolicyWindow.hitRatio = win ? (float)((double)persist / (double)win > 1.0 ? 1.0 : (double)persist / (double)win) : 0.f;
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  CU(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
  return HSPMV_OK;
}

template <typename T>
int dispatch(hspmv_matrix *A, int algo, const void *dx, void *dy, cudaStream_t st, int64_t row_lo = 0,
             int64_t row_hi = -1) {
  const T *x = (const T *)(dx ? dx : A->d_x);
  T *y = (T *)(dy ? dy : A->d_y);
  if (A->m == 0) return HSPMV_OK;
  if (row_hi < 0) row_hi = A->m;
  const bool ranged = row_lo != 0 || row_hi != A->m;
  if (ranged) {
    if (row_lo < 0 || row_lo > row_hi || row_hi > A->m) return fail(HSPMV_EINVAL, "bad row range");
    if (row_lo == row_hi) return HSPMV_OK;
    if (algo == HSPMV_ALGO_AUTO) algo = HSPMV_ALGO_CSR_STREAM;
    if (algo != HSPMV_ALGO_CSR_STREAM && algo != HSPMV_ALGO_CSR_VECTOR)
      return fail(HSPMV_EUNSUPPORTED, "row ranges are supported by CSR_STREAM and CSR_VECTOR only");
  }
  int rc = apply_l2_window(A, x, st);
  if (rc) return rc;
  if (algo == HSPMV_ALGO_AUTO) {
    if (A->k >= 2) {
      algo = HSPMV_ALGO_CSRK;
    } else {
      rc = ensure_stats(A);
      if (rc) return rc;
      const double mean = A->m ? (double)A->nnz / (double)A->m : 0.0;
      // regular rows: the TMA tile engine.  Skewed row lengths (power law): the row-length-binned
      // vector kernels -- on R-MAT scale 24 they run at 2.06 TB/s against 0.67 TB/s for the merge
      // kernel (profiles/r1_configs_1to4_fullsize.jsonl); the random x gather (32 B sector per 8 B
      // used) caps that matrix near 0.4 of the roofline whatever the row split.
      algo = (A->stats[8] <= 8 * (mean + 1) + 64) ? HSPMV_ALGO_CSR_STREAM : HSPMV_ALGO_CSR_VECTOR;
      // fp32, short or ragged rows: the streamed bytes halve but the TMA tile engine's per-tile instruction cost does
      // not, and the row-block kernel (products staged in shared memory, occupancy instead of a pipeline) is 8-25 %
      // faster (profiles/r2_sweep_block_cfg1to4.jsonl: cfg1 15.3 vs 17.3 us, cfg3 95 vs 120 us; fp64 and the 27-point
      // stencil stay with the tile engine)
      if (algo == HSPMV_ALGO_CSR_STREAM && !ranged && sizeof(T) == 4 && (mean < 8.0 || (double)A->stats[8] > mean * 1.15 + 1.0))
        algo = HSPMV_ALGO_CSR_BLOCK;
    }
  }
  switch (algo) {
  case HSPMV_ALGO_CSR_VECTOR: rc = run_csr_vector<T>(A, x, y, st, row_lo, row_hi); break;
  case HSPMV_ALGO_CSR_STREAM: rc = run_csrk<T>(A, true, x, y, st, row_lo, row_hi); break;
  case HSPMV_ALGO_CSR_MERGE: rc = run_csr_merge<T>(A, x, y, st); break;
  case HSPMV_ALGO_CSR_BLOCK: rc = run_csr_block<T>(A, x, y, st); break;
  case HSPMV_ALGO_CSRK: rc = run_csrk<T>(A, false, x, y, st); break;
  default: return fail(HSPMV_EINVAL, "unknown algo");
  }
  if (rc) return rc;
  CU(cudaGetLastError());
  return HSPMV_OK;
}

int do_spmv(hspmv_matrix *A, int algo, const void *dx, void *dy, cudaStream_t st, int64_t row_lo = 0,
            int64_t row_hi = -1) {
  A->last_stream = st;
  A->last_stream_valid = true;
  return A->dtype == HSPMV_F64 ? dispatch<double>(A, algo, dx, dy, st, row_lo, row_hi)
                               : dispatch<float>(A, algo, dx, dy, st, row_lo, row_hi);
}

int upload_csrk_maps(hspmv_matrix *A) {
  A->sr_max_rows = A->sr_max_nnz = 0;
  if (!A->h_rp.empty())
    for (int64_t s = 0; s < A->n_sr; ++s) {
      const int64_t r0 = A->h_sr[(size_t)s], r1 = A->h_sr[(size_t)s + 1];
      A->sr_max_rows = std::max(A->sr_max_rows, r1 - r0);
      A->sr_max_nnz = std::max(A->sr_max_nnz, (int64_t)A->h_rp[(size_t)r1] - A->h_rp[(size_t)r0]);
    }
  for (u32 **p : {&A->d_sr, &A->d_ssr, &A->d_sr_e, &A->d_perm})
    if (*p) {
      cudaFree(*p);
      *p = nullptr;
    }
  CU(cudaMalloc(&A->d_sr, (size_t)(A->n_sr + 1 + PAD) * 4));
  CU(cudaMemset(A->d_sr, 0, (size_t)(A->n_sr + 1 + PAD) * 4));
  CU(cudaMemcpy(A->d_sr, A->h_sr.data(), (size_t)(A->n_sr + 1) * 4, cudaMemcpyHostToDevice));
  CU(cudaMalloc(&A->d_sr_e, (size_t)(A->n_sr + 1 + PAD) * 4));
  hspmv::sr_offsets_kernel<<<grid_for(A->n_sr + 1, 256, A->dev.sms * 8), 256, 0, A->stream>>>(A->n_sr, A->d_sr, A->d_rp, A->d_sr_e);
  LAUNCHED();
  if (A->k == 3) {
    CU(cudaMalloc(&A->d_ssr, (size_t)(A->n_ssr + 1 + PAD) * 4));
    CU(cudaMemset(A->d_ssr, 0, (size_t)(A->n_ssr + 1 + PAD) * 4));
    CU(cudaMemcpy(A->d_ssr, A->h_ssr.data(), (size_t)(A->n_ssr + 1) * 4, cudaMemcpyHostToDevice));
  }
  if (!A->h_perm.empty()) {
    CU(cudaMalloc(&A->d_perm, (size_t)(A->m + PAD) * 4));
    CU(cudaMemcpy(A->d_perm, A->h_perm.data(), (size_t)A->m * 4, cudaMemcpyHostToDevice));
  }
  CU(cudaStreamSynchronize(A->stream));
  return HSPMV_OK;
}

} // namespace

// =============================================================================================
extern "C" {

int hspmv_version(void) { return 100; }
const char *hspmv_last_error(void) { return g_err.c_str(); }

int hspmv_device_count(int *count) {
  if (!count) return fail(HSPMV_EINVAL, "null");
  CU(cudaGetDeviceCount(count));
  return HSPMV_OK;
}
int hspmv_set_device(int device) {
  CU(cudaSetDevice(device));
  return HSPMV_OK;
}
int hspmv_device_info(int *sm_count, int64_t *l2_bytes, int64_t *hbm_bytes, int *cc) {
  DeviceProps p;
  int rc = query_props(p);
  if (rc) return rc;
  if (sm_count) *sm_count = p.sms;
  if (l2_bytes) *l2_bytes = p.l2;
  if (hbm_bytes) *hbm_bytes = p.hbm;
  if (cc) *cc = p.cc;
  return HSPMV_OK;
}
int64_t hspmv_launch_count(void) { return g_launches.load(); }
int hspmv_omp_spmv_cache_size(void);

// ---- files ----------------------------------------------------------------------------
int hspmv_read_csr_text(const char *path, int index_base, int dtype, int64_t *m, int64_t *n,
                        int64_t *nnz, int32_t **row_ptr, int32_t **col_idx, void **val) {
  std::string err;
  int rc = hspmv::read_csr_text(path, index_base, dtype, m, n, nnz, row_ptr, col_idx, val, err);
  return rc ? fail(rc, err) : rc;
}
int hspmv_read_csr_cached(const char *path, int index_base, int dtype, int64_t *m, int64_t *n, int64_t *nnz,
                          int32_t **row_ptr, int32_t **col_idx, void **val, int *cache_hit) {
  std::string err;
  int rc = hspmv::read_csr_cached(path, index_base, dtype, m, n, nnz, row_ptr, col_idx, val, cache_hit, err);
  return rc ? fail(rc, err.empty() ? "read_csr_cached: bad arguments" : err) : rc;
}
int hspmv_write_csr_text(const char *path, int index_base, int dtype, int64_t m, int64_t n,
                         int64_t nnz, const int32_t *row_ptr, const int32_t *col_idx, const void *val) {
  std::string err;
  int rc = hspmv::write_csr_text(path, index_base, dtype, m, n, nnz, row_ptr, col_idx, val, err);
  return rc ? fail(rc, err) : rc;
}
int hspmv_read_csr3_text(const char *path, int dtype, int64_t *n_ssr, int64_t *n_sr, int64_t *m,
                         int64_t *n, int64_t *nnz, uint32_t **ssr_ptr, uint32_t **sr_ptr,
                         uint32_t **row_ptr, uint32_t **col_idx, void **val) {
  std::string err;
  int rc = hspmv::read_csr3_text(path, dtype, n_ssr, n_sr, m, n, nnz, ssr_ptr, sr_ptr, row_ptr, col_idx, val, err);
  return rc ? fail(rc, err) : rc;
}
int hspmv_write_csr3_text(const char *path, int dtype, int64_t n_ssr, int64_t n_sr, int64_t m,
                          int64_t n, int64_t nnz, const uint32_t *ssr_ptr, const uint32_t *sr_ptr,
                          const uint32_t *row_ptr, const uint32_t *col_idx, const void *val) {
  std::string err;
  int rc = hspmv::write_csr3_text(path, dtype, n_ssr, n_sr, m, n, nnz, ssr_ptr, sr_ptr, row_ptr, col_idx, val, err);
  return rc ? fail(rc, err) : rc;
}
void hspmv_free_host(void *p) { std::free(p); }

// ---- handles --------------------------------------------------------------------------
int hspmv_create_csr(hspmv_matrix **out, int64_t m, int64_t n, int64_t nnz, const int32_t *row_ptr,
                     const int32_t *col_idx, const void *val, int dtype, int index_base, int keep_host) {
  if (!row_ptr || (nnz > 0 && (!col_idx || !val))) return fail(HSPMV_EINVAL, "null CSR array");
  if ((int64_t)row_ptr[m] - index_base != nnz) return fail(HSPMV_EINVAL, "row_ptr[m] != nnz");
  hspmv_matrix *A = nullptr;
  int rc = new_handle(&A, dtype, m, n, nnz);
  if (rc) return rc;
  const size_t vs = A->vsize();
  // host copy (also the place where a 1-based file becomes 0-based)
  A->h_rp.resize((size_t)m + 1);
  A->h_ci.resize((size_t)nnz);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i <= m; ++i) A->h_rp[(size_t)i] = row_ptr[i] - index_base;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nnz; ++i) A->h_ci[(size_t)i] = col_idx[i] - index_base;
  for (int64_t i = 0; i < m; ++i)
    if (A->h_rp[(size_t)i] > A->h_rp[(size_t)i + 1] || A->h_rp[(size_t)i] < 0) {
      hspmv_destroy(A);
      return fail(HSPMV_EINVAL, "row_ptr is not monotone");
    }
  bool bad_col = false;
#pragma omp parallel for schedule(static) reduction(|| : bad_col)
  for (int64_t i = 0; i < nnz; ++i) bad_col = bad_col || A->h_ci[(size_t)i] < 0 || A->h_ci[(size_t)i] >= n;
  if (bad_col) {
    hspmv_destroy(A);
    return fail(HSPMV_EINVAL, "column index out of range (wrong index_base?)");
  }
  if ((rc = alloc_matrix_arrays(A)) || (rc = alloc_vectors(A))) {
    hspmv_destroy(A);
    return rc;
  }
  cudaError_t e = cudaMemcpy(A->d_rp, A->h_rp.data(), (size_t)(m + 1) * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nnz) e = cudaMemcpy(A->d_ci, A->h_ci.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nnz) e = cudaMemcpy(A->d_val, val, (size_t)nnz * vs, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    hspmv_destroy(A);
    return fail(HSPMV_ECUDA, std::string("upload: ") + cudaGetErrorString(e));
  }
  if (keep_host) {
    A->h_val.assign((const unsigned char *)val, (const unsigned char *)val + (size_t)nnz * vs);
  } else {
    ivec().swap(A->h_ci);
  }
  hspmv::row_stats(m, A->h_rp.data(), A->stats);
  A->stats_ready = true;
  *out = A;
  return HSPMV_OK;
}

int hspmv_create_csrk(hspmv_matrix **out, int64_t n_ssr, int64_t n_sr, int64_t m, int64_t n,
                      int64_t nnz, const uint32_t *ssr_ptr, const uint32_t *sr_ptr,
                      const uint32_t *row_ptr, const uint32_t *col_idx, const void *val, int dtype) {
  if (!sr_ptr || n_sr <= 0) return fail(HSPMV_EINVAL, "sr_ptr required");
  if (sr_ptr[0] != 0 || sr_ptr[n_sr] != (uint32_t)m) return fail(HSPMV_EINVAL, "sr_ptr does not cover the rows");
  if (ssr_ptr && (ssr_ptr[0] != 0 || ssr_ptr[n_ssr] != (uint32_t)n_sr)) return fail(HSPMV_EINVAL, "ssr_ptr does not cover the super-rows");
  hspmv_matrix *A = nullptr;
  int rc = hspmv_create_csr(&A, m, n, nnz, (const int32_t *)row_ptr, (const int32_t *)col_idx, val, dtype, 0, 0);
  if (rc) return rc;
  A->k = ssr_ptr ? 3 : 2;
  A->n_sr = n_sr;
  A->n_ssr = ssr_ptr ? n_ssr : 0;
  A->h_sr.assign(sr_ptr, sr_ptr + n_sr + 1);
  if (ssr_ptr) A->h_ssr.assign(ssr_ptr, ssr_ptr + n_ssr + 1);
  rc = upload_csrk_maps(A); // no permutation known: x / y are taken in the file's (reordered) order
  if (rc) {
    hspmv_destroy(A);
    return rc;
  }
  *out = A;
  return HSPMV_OK;
}

int hspmv_create_stencil(hspmv_matrix **out, int kind, int64_t g, int dtype, int mode, int64_t row_lo,
                         int64_t row_hi, int64_t col_shift, int64_t n_cols) {
  if ((kind != 27 && kind != 5) || g < 1) return fail(HSPMV_EINVAL, "kind must be 27 or 5");
  const int64_t n_all = kind == 27 ? g * g * g : g * g;
  if (row_hi < 0) row_hi = n_all;
  if (row_lo < 0 || row_lo > row_hi || row_hi > n_all) return fail(HSPMV_EINVAL, "bad row range");
  if (n_cols <= 0) n_cols = n_all - col_shift;
  const int64_t before_lo = kind == 27 ? hspmv::stencil27_before(row_lo, g) : hspmv::laplace5_before(row_lo, g);
  const int64_t before_hi = kind == 27 ? hspmv::stencil27_before(row_hi, g) : hspmv::laplace5_before(row_hi, g);
  hspmv_matrix *A = nullptr;
  int rc = new_handle(&A, dtype, row_hi - row_lo, n_cols, before_hi - before_lo);
  if (rc) return rc;
  if ((rc = alloc_matrix_arrays(A)) || (rc = alloc_vectors(A))) {
    hspmv_destroy(A);
    return rc;
  }
  const int grid = grid_for(A->m + 1, 256, A->dev.sms * 16);
  if (kind == 27) {
    if (dtype == HSPMV_F64)
      hspmv::stencil27_kernel<double><<<grid, 256, 0, A->stream>>>(g, row_lo, row_hi, col_shift, mode, A->d_rp, A->d_ci, (double *)A->d_val);
    else
      hspmv::stencil27_kernel<float><<<grid, 256, 0, A->stream>>>(g, row_lo, row_hi, col_shift, mode, A->d_rp, A->d_ci, (float *)A->d_val);
  } else {
    if (dtype == HSPMV_F64)
      hspmv::laplace5_kernel<double><<<grid, 256, 0, A->stream>>>(g, row_lo, row_hi, col_shift, mode, A->d_rp, A->d_ci, (double *)A->d_val);
    else
      hspmv::laplace5_kernel<float><<<grid, 256, 0, A->stream>>>(g, row_lo, row_hi, col_shift, mode, A->d_rp, A->d_ci, (float *)A->d_val);
  }
  LAUNCHED();
  cudaError_t e = cudaStreamSynchronize(A->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    hspmv_destroy(A);
    return fail(HSPMV_ECUDA, std::string("stencil generator: ") + cudaGetErrorString(e));
  }
  // closed-form statistics (no host copy of row_ptr for the 64 M-row configuration)
  for (int i = 0; i < 10; ++i) A->stats[i] = 0;
  A->stats[8] = kind == 27 ? (g >= 3 ? 27 : g == 2 ? 8 : 1) : (g >= 3 ? 5 : g == 2 ? 3 : 1);
  A->stats[hspmv::bin_of_len(A->stats[8])] = A->m; // every row is within the same vector bin class
  A->stats_ready = true;
  *out = A;
  return HSPMV_OK;
}

int hspmv_destroy(hspmv_matrix *A) {
  if (!A) return HSPMV_OK;
  free_device(A);
  if (A->stream) cudaStreamDestroy(A->stream);
  delete A;
  return HSPMV_OK;
}

int hspmv_matrix_info(const hspmv_matrix *A, int64_t *m, int64_t *n, int64_t *nnz, int *dtype, int *k,
                      int64_t *n_sr, int64_t *n_ssr) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  if (m) *m = A->m;
  if (n) *n = A->n;
  if (nnz) *nnz = A->nnz;
  if (dtype) *dtype = A->dtype;
  if (k) *k = A->k;
  if (n_sr) *n_sr = A->n_sr;
  if (n_ssr) *n_ssr = A->n_ssr;
  return HSPMV_OK;
}

int hspmv_download_csr(const hspmv_matrix *A, int32_t *row_ptr, int32_t *col_idx, void *val) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  if (row_ptr) CU(cudaMemcpy(row_ptr, A->d_rp, (size_t)(A->m + 1) * 4, cudaMemcpyDeviceToHost));
  if (col_idx && A->nnz) CU(cudaMemcpy(col_idx, A->d_ci, (size_t)A->nnz * 4, cudaMemcpyDeviceToHost));
  if (val && A->nnz) CU(cudaMemcpy(val, A->d_val, (size_t)A->nnz * A->vsize(), cudaMemcpyDeviceToHost));
  return HSPMV_OK;
}

int hspmv_row_stats(const hspmv_matrix *A, int64_t out[10]) {
  if (!A || !out) return fail(HSPMV_EINVAL, "null");
  int rc = ensure_stats(const_cast<hspmv_matrix *>(A));
  if (rc) return rc;
  for (int i = 0; i < 10; ++i) out[i] = A->stats[i];
  return HSPMV_OK;
}

// ---- Band-k -----------------------------------------------------------------------------
int hspmv_bandk_host(int64_t n, int64_t nnz, const uint32_t *row_ptr, const uint32_t *col_idx, int k,
                     const int *sup_row_sizes, uint32_t *perm, uint32_t *sr_ptr, int64_t *n_sr,
                     uint32_t *ssr_ptr, int64_t *n_ssr) {
  uvec p, sr, ssr;
  int rc = hspmv::bandk_build(n, nnz, row_ptr, col_idx, k, sup_row_sizes, p, sr, ssr);
  if (rc) return fail(rc, rc == HSPMV_EUNSUPPORTED
                              ? "Band-k: a coarse vertex has no upper/diagonal entry (matrix must be structurally symmetric with a stored diagonal)"
                              : "Band-k: bad arguments");
  if (perm) std::memcpy(perm, p.data(), p.size() * 4);
  if (sr_ptr) std::memcpy(sr_ptr, sr.data(), sr.size() * 4);
  if (n_sr) *n_sr = (int64_t)sr.size() - 1;
  if (k == 3 && ssr_ptr) std::memcpy(ssr_ptr, ssr.data(), ssr.size() * 4);
  if (n_ssr) *n_ssr = k == 3 ? (int64_t)ssr.size() - 1 : 0;
  return HSPMV_OK;
}

int hspmv_reorder_host(int64_t n, int64_t nnz, const uint32_t *perm, const uint32_t *row_ptr,
                       const uint32_t *col_idx, const void *val, int dtype, uint32_t *row_ptr_out,
                       uint32_t *col_idx_out, void *val_out) {
  (void)nnz;
  if (!perm || !row_ptr || !col_idx || !val || !row_ptr_out || !col_idx_out || !val_out)
    return fail(HSPMV_EINVAL, "null");
  if (dtype == HSPMV_F64)
    hspmv::reorder_f64(n, perm, row_ptr, col_idx, (const double *)val, row_ptr_out, col_idx_out, (double *)val_out);
  else
    hspmv::reorder_f32(n, perm, row_ptr, col_idx, (const float *)val, row_ptr_out, col_idx_out, (float *)val_out);
  return HSPMV_OK;
}

} // extern "C"

namespace {

// CSR-k bookkeeping shared by the host and the device builder: maps, stats, stale caches
int finish_csrk(hspmv_matrix *A, int k, uvec &perm, uvec &sr, uvec &ssr) {
  A->k = k;
  A->h_perm.swap(perm);
  A->h_sr.swap(sr);
  A->h_ssr.swap(ssr);
  A->n_sr = (int64_t)A->h_sr.size() - 1;
  A->n_ssr = k == 3 ? (int64_t)A->h_ssr.size() - 1 : 0;
  A->bins_ready = false;
  A->merge_tile = 0;
  A->block_cap = 0;
  A->graph_iters = 0; // a captured graph points at the old arrays
  free_pipe(A);
  hspmv::row_stats(A->m, A->h_rp.data(), A->stats);
  A->stats_ready = true;
  return upload_csrk_maps(A);
}

const char *bandk_error(int rc) {
  return rc == HSPMV_EUNSUPPORTED
             ? "Band-k: a coarse vertex has no upper/diagonal entry (matrix must be structurally symmetric with a stored diagonal)"
             : "Band-k: bad arguments (k must be 2 or 3)";
}

// Device builder: the O(nnz) steps (level-1 coarse graph, A' = P A P^T) run in HBM on the
// resident arrays; only row_ptr, the coarse graph and the permutation cross PCIe.
// Returns BANDK_USE_HOST when the host builder has to take over.
int build_csrk_device(hspmv_matrix *A, int k, const int *sizes) {
  const size_t vs = A->vsize();
  typedef std::chrono::steady_clock Clock;
  Clock::time_point t_prev = Clock::now();
  const Clock::time_point t_begin = t_prev;
  for (double &d : A->build_phase) d = 0;
  auto lap = [&](int phase) {
    const Clock::time_point now = Clock::now();
    A->build_phase[phase] += std::chrono::duration<double>(now - t_prev).count();
    t_prev = now;
  };
  if (A->h_rp.empty()) {
    A->h_rp.resize((size_t)A->m + 1);
    hspmv::first_touch(A->h_rp.data(), (size_t)(A->m + 1) * 4);
    CU(cudaMemcpy(A->h_rp.data(), A->d_rp, (size_t)(A->m + 1) * 4, cudaMemcpyDeviceToHost));
  }
  if (A->nnz == 0 || A->h_rp[(size_t)A->m] != (int32_t)A->nnz) return hspmv::BANDK_USE_HOST;
  int rc0 = ensure_stats(A);
  if (rc0) return rc0;
  if (A->stats[8] > hspmv::BANDK_MAX_ROW) return hspmv::BANDK_USE_HOST; // longest row vs the warp sorter
  uvec start;
  hspmv::bandk_group_rows((const u32 *)A->h_rp.data(), A->m, hspmv::bandk_threshold(sizes[0], A->nnz, A->m), start);
  if (start.back() != (u32)A->m) return hspmv::BANDK_USE_HOST; // trailing empty rows: the host path reports it
  lap(0);
  hspmv::CoarseGraph level1;
  int64_t launches = 0;
  int rc = hspmv::coarse_level1_device(A->m, A->nnz, (const u32 *)A->d_rp, (const u32 *)A->d_ci, start, A->dev.sms,
                                       A->stream, level1, &launches);
  g_launches.fetch_add(launches, std::memory_order_relaxed);
  if (rc == hspmv::BANDK_USE_HOST) return rc;
  if (rc == HSPMV_ECUDA) return fail(rc, std::string("Band-k device build: ") + cudaGetErrorString(cudaGetLastError()));
  if (rc) return fail(rc, bandk_error(rc));
  lap(1);
  uvec perm, sr, ssr;
  rc = hspmv::bandk_finish(A->m, k, sizes, level1, start, perm, sr, ssr);
  if (rc) return fail(rc, bandk_error(rc));
  lap(2);

  // reordered arrays next to the old ones, then swap
  int *rp2 = nullptr, *ci2 = nullptr;
  void *val2 = nullptr;
  u32 *d_perm = nullptr;
  auto drop = [&]() {
    cudaFree(rp2);
    cudaFree(ci2);
    cudaFree(val2);
    cudaFree(d_perm);
    cudaGetLastError();
  };
  if (cudaMalloc(&rp2, (size_t)(A->m + 1 + PAD) * 4) != cudaSuccess || cudaMalloc(&ci2, (size_t)(A->nnz + PAD) * 4) != cudaSuccess ||
      cudaMalloc(&val2, (size_t)(A->nnz + PAD) * vs) != cudaSuccess || cudaMalloc(&d_perm, (size_t)A->m * 4) != cudaSuccess) {
    drop();
    return hspmv::BANDK_USE_HOST;
  }
  cudaMemsetAsync((char *)rp2 + (size_t)(A->m + 1) * 4, 0, PAD * 4, A->stream);
  cudaMemsetAsync((char *)ci2 + (size_t)A->nnz * 4, 0, PAD * 4, A->stream);
  cudaMemsetAsync((char *)val2 + (size_t)A->nnz * vs, 0, PAD * vs, A->stream);
  cudaMemcpyAsync(d_perm, perm.data(), (size_t)A->m * 4, cudaMemcpyHostToDevice, A->stream);
  launches = 0;
  rc = hspmv::reorder_device(A->dtype, A->m, A->nnz, d_perm, (const u32 *)A->d_rp, (const u32 *)A->d_ci, A->d_val,
                             (u32 *)rp2, (u32 *)ci2, val2, A->dev.sms, A->stream, &launches);
  g_launches.fetch_add(launches, std::memory_order_relaxed);
  if (rc) {
    drop();
    if (rc == hspmv::BANDK_USE_HOST) return rc;
    return fail(rc, std::string("Band-k device reorder: ") + cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(d_perm);
  cudaFree(A->d_rp);
  cudaFree(A->d_ci);
  cudaFree(A->d_val);
  A->d_rp = rp2;
  A->d_ci = ci2;
  A->d_val = val2;
  CU(cudaMemcpy(A->h_rp.data(), A->d_rp, (size_t)(A->m + 1) * 4, cudaMemcpyDeviceToHost));
  ivec().swap(A->h_ci); // host copies of the un-reordered entries are stale now
  std::vector<unsigned char>().swap(A->h_val);
  lap(3);
  rc = finish_csrk(A, k, perm, sr, ssr);
  lap(4);
  A->build_phase[5] = std::chrono::duration<double>(Clock::now() - t_begin).count();
  return rc;
}

} // namespace

extern "C" {

int hspmv_build_csrk(hspmv_matrix *A, int k, const int *sup_row_sizes) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  if (A->m != A->n) return fail(HSPMV_EINVAL, "Band-k needs a square matrix");
  if (A->k != 1) return fail(HSPMV_ESTATE, "matrix is already in CSR-k form");
  if (k < 2 || k > 3 || !sup_row_sizes || A->m <= 0) return fail(HSPMV_EINVAL, bandk_error(HSPMV_EINVAL));
  if (!A->opt[HSPMV_OPT_BANDK_HOST]) {
    int rc = build_csrk_device(A, k, sup_row_sizes);
    if (rc != hspmv::BANDK_USE_HOST) return rc;
  }
  const size_t vs = A->vsize();
  // host copy of the CSR arrays: kept at creation, or fetched back from HBM
  if (A->h_rp.empty()) {
    A->h_rp.resize((size_t)A->m + 1);
    hspmv::first_touch(A->h_rp.data(), (size_t)(A->m + 1) * 4);
    CU(cudaMemcpy(A->h_rp.data(), A->d_rp, (size_t)(A->m + 1) * 4, cudaMemcpyDeviceToHost));
  }
  if (A->h_ci.size() != (size_t)A->nnz) {
    A->h_ci.resize((size_t)A->nnz);
    hspmv::first_touch(A->h_ci.data(), (size_t)A->nnz * 4);
    if (A->nnz) CU(cudaMemcpy(A->h_ci.data(), A->d_ci, (size_t)A->nnz * 4, cudaMemcpyDeviceToHost));
  }
  if (A->h_val.size() != (size_t)A->nnz * vs) {
    A->h_val.resize((size_t)A->nnz * vs);
    if (A->nnz) CU(cudaMemcpy(A->h_val.data(), A->d_val, (size_t)A->nnz * vs, cudaMemcpyDeviceToHost));
  }
  uvec perm, sr, ssr;
  int rc = hspmv::bandk_build(A->m, A->nnz, (const u32 *)A->h_rp.data(), (const u32 *)A->h_ci.data(), k,
                              sup_row_sizes, perm, sr, ssr);
  if (rc) return fail(rc, bandk_error(rc));
  if (sr.back() != (u32)A->m) return fail(HSPMV_EUNSUPPORTED, "Band-k: trailing empty rows are not covered by a super-row");
  // A' = P A P^T on the host, then one upload of the REORDERED arrays
  ivec rp2((size_t)A->m + 1), ci2((size_t)A->nnz);
  std::vector<unsigned char> val2((size_t)A->nnz * vs);
  if (A->dtype == HSPMV_F64)
    hspmv::reorder_f64(A->m, perm.data(), (const u32 *)A->h_rp.data(), (const u32 *)A->h_ci.data(),
                       (const double *)A->h_val.data(), (u32 *)rp2.data(), (u32 *)ci2.data(), (double *)val2.data());
  else
    hspmv::reorder_f32(A->m, perm.data(), (const u32 *)A->h_rp.data(), (const u32 *)A->h_ci.data(),
                       (const float *)A->h_val.data(), (u32 *)rp2.data(), (u32 *)ci2.data(), (float *)val2.data());
  A->h_rp.swap(rp2);
  A->h_ci.swap(ci2);
  A->h_val.swap(val2);
  CU(cudaMemcpy(A->d_rp, A->h_rp.data(), (size_t)(A->m + 1) * 4, cudaMemcpyHostToDevice));
  if (A->nnz) {
    CU(cudaMemcpy(A->d_ci, A->h_ci.data(), (size_t)A->nnz * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(A->d_val, A->h_val.data(), (size_t)A->nnz * vs, cudaMemcpyHostToDevice));
  }
  return finish_csrk(A, k, perm, sr, ssr);
}

// Power-law matrices: renumber the COLUMNS by how often they are referenced, most popular first.  The x entries that
// the gathers hit most become neighbours in memory (four of them per 32-byte sector, all hot), so the gathers hit L1 /
// L2 instead of pulling a sector per entry from all over a 134 MB vector.  Rows are not moved: y keeps its numbering;
// hspmv_set_x / hspmv_spmv_host permute x on the way in (x_kernel[i] = x[perm[i]]).
int hspmv_relabel_hub_columns(hspmv_matrix *A, uint32_t *perm_out) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  if (A->k != 1 || A->d_perm || A->sts_ready) return fail(HSPMV_ESTATE, "matrix is already reordered (CSR-k, STS or hub columns)");
  if (A->n <= 0) return HSPMV_OK;
  unsigned int *d_cnt = nullptr;
  CU(cudaMalloc(&d_cnt, (size_t)A->n * 4));
  CU(cudaMemsetAsync(d_cnt, 0, (size_t)A->n * 4, A->stream));
  if (A->nnz) {
    hspmv::col_count_kernel<<<grid_for(A->nnz, 256, A->dev.sms * 16), 256, 0, A->stream>>>(A->nnz, A->d_ci, d_cnt);
    LAUNCHED();
  }
  uvec cnt((size_t)A->n), perm, inv;
  hspmv::first_touch(cnt.data(), (size_t)A->n * 4);
  CU(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)A->n * 4, cudaMemcpyDeviceToHost, A->stream));
  CU(cudaStreamSynchronize(A->stream));
  // counting sort by descending count, ascending column id inside a count (deterministic; partition.cpp: hub_order)
  hspmv::hub_order(A->n, cnt.data(), perm, inv);
  CU(cudaMemcpyAsync(d_cnt, inv.data(), (size_t)A->n * 4, cudaMemcpyHostToDevice, A->stream)); // reuse as new_id[]
  if (A->nnz) {
    hspmv::col_relabel_kernel<<<grid_for(A->nnz, 256, A->dev.sms * 16), 256, 0, A->stream>>>(A->nnz, d_cnt, A->d_ci);
    LAUNCHED();
  }
  CU(cudaStreamSynchronize(A->stream));
  cudaFree(d_cnt);
  CU(cudaMalloc(&A->d_perm, (size_t)(A->n + PAD) * 4));
  CU(cudaMemcpy(A->d_perm, perm.data(), (size_t)A->n * 4, cudaMemcpyHostToDevice));
  A->cols_only_perm = true;
  A->graph_iters = 0; // (the arrays stay where they are, but a replay must not be mistaken for the un-relabelled product)
  ivec().swap(A->h_ci); // host copy (if any) is stale
  free_pipe(A);
  if (perm_out) std::memcpy(perm_out, perm.data(), perm.size() * 4);
  A->h_perm.swap(perm);
  return HSPMV_OK;
}

int hspmv_csrk_get(const hspmv_matrix *A, uint32_t *perm, uint32_t *sr_ptr, uint32_t *ssr_ptr) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  if (A->k < 2) return fail(HSPMV_ESTATE, "not in CSR-k form");
  if (perm) {
    if (A->h_perm.empty()) return fail(HSPMV_ESTATE, "no permutation (matrix came from a .csr3 file)");
    std::memcpy(perm, A->h_perm.data(), A->h_perm.size() * 4);
  }
  if (sr_ptr) std::memcpy(sr_ptr, A->h_sr.data(), A->h_sr.size() * 4);
  if (ssr_ptr && A->k == 3) std::memcpy(ssr_ptr, A->h_ssr.data(), A->h_ssr.size() * 4);
  return HSPMV_OK;
}

int hspmv_auto_params(int64_t nnz, int64_t n_rows, int arch, int sizes[2], int *veclevel, int *block_dim_y) {
  int rc = hspmv::auto_params(nnz, n_rows, arch, sizes, veclevel, block_dim_y);
  return rc ? fail(rc, "auto_params: bad arguments") : rc;
}

// ---- vectors ----------------------------------------------------------------------------
int hspmv_set_x(hspmv_matrix *A, const void *x_host) {
  if (!A || !x_host) return fail(HSPMV_EINVAL, "null");
  const size_t vs = A->vsize();
  if (A->d_perm) {
    CU(cudaMemcpyAsync(A->d_tmp, x_host, (size_t)A->n * vs, cudaMemcpyHostToDevice, A->stream));
    const int grid = grid_for(A->n, 256, A->dev.sms * 16);
    if (A->dtype == HSPMV_F64)
      hspmv::gather_kernel<double><<<grid, 256, 0, A->stream>>>(A->n, A->d_perm, (const double *)A->d_tmp, (double *)A->d_x);
    else
      hspmv::gather_kernel<float><<<grid, 256, 0, A->stream>>>(A->n, A->d_perm, (const float *)A->d_tmp, (float *)A->d_x);
    LAUNCHED();
  } else {
    CU(cudaMemcpyAsync(A->d_x, x_host, (size_t)A->n * vs, cudaMemcpyHostToDevice, A->stream));
  }
  CU(cudaStreamSynchronize(A->stream));
  return HSPMV_OK;
}

int hspmv_get_y(hspmv_matrix *A, void *y_host, int original_order) {
  if (!A || !y_host) return fail(HSPMV_EINVAL, "null");
  const size_t vs = A->vsize();
  // the last SpMV may have been enqueued on the caller's stream: order the read-back after it
  if (A->last_stream_valid && A->last_stream != A->stream) CU(cudaStreamSynchronize(A->last_stream));
  if (original_order && A->d_perm && !A->cols_only_perm) {
    const int grid = grid_for(A->m, 256, A->dev.sms * 16);
    if (A->dtype == HSPMV_F64)
      hspmv::scatter_kernel<double><<<grid, 256, 0, A->stream>>>(A->m, A->d_perm, (const double *)A->d_y, (double *)A->d_tmp);
    else
      hspmv::scatter_kernel<float><<<grid, 256, 0, A->stream>>>(A->m, A->d_perm, (const float *)A->d_y, (float *)A->d_tmp);
    LAUNCHED();
    CU(cudaMemcpyAsync(y_host, A->d_tmp, (size_t)A->m * vs, cudaMemcpyDeviceToHost, A->stream));
  } else {
    CU(cudaMemcpyAsync(y_host, A->d_y, (size_t)A->m * vs, cudaMemcpyDeviceToHost, A->stream));
  }
  CU(cudaStreamSynchronize(A->stream));
  return HSPMV_OK;
}

int hspmv_set_y(hspmv_matrix *A, const void *y_host) {
  if (!A || !y_host) return fail(HSPMV_EINVAL, "null");
  CU(cudaMemcpyAsync(A->d_y, y_host, (size_t)A->m * A->vsize(), cudaMemcpyHostToDevice, A->stream));
  CU(cudaStreamSynchronize(A->stream));
  return HSPMV_OK;
}

int hspmv_sync(hspmv_matrix *A) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  if (A->last_stream_valid) CU(cudaStreamSynchronize(A->last_stream));
  CU(cudaStreamSynchronize(A->stream));
  return HSPMV_OK;
}

int hspmv_build_phases(const hspmv_matrix *A, double *seconds, int n) {
  if (!A || !seconds || n < 0) return fail(HSPMV_EINVAL, "null");
  for (int i = 0; i < n; ++i) seconds[i] = i < 6 ? A->build_phase[i] : 0.0;
  return HSPMV_OK;
}

int hspmv_device_view_get(hspmv_matrix *A, hspmv_device_view *out) {
  if (!A || !out) return fail(HSPMV_EINVAL, "null");
  const int64_t coarsest = A->k == 3 ? A->n_ssr : A->k == 2 ? A->n_sr : A->m;
  if (!A->d_dims) CU(cudaMalloc(&A->d_dims, 4 * sizeof(u32)));
  const u32 dims[4] = {(u32)A->m, (u32)A->n, (u32)A->nnz, (u32)coarsest};
  CU(cudaMemcpy(A->d_dims, dims, sizeof(dims), cudaMemcpyHostToDevice));
  out->row_ptr = A->d_rp;
  out->col_idx = A->d_ci;
  out->val = A->d_val;
  out->x = A->d_x;
  out->y = A->d_y;
  out->sr_ptr = A->k >= 2 ? A->d_sr : nullptr;
  out->ssr_ptr = A->k == 3 ? A->d_ssr : nullptr;
  out->dims = A->d_dims;
  out->n_coarsest = coarsest;
  out->dtype = A->dtype;
  out->k = A->k;
  return HSPMV_OK;
}

void *hspmv_x_device(hspmv_matrix *A) { return A ? A->d_x : nullptr; }
void *hspmv_y_device(hspmv_matrix *A) { return A ? A->d_y : nullptr; }

int hspmv_fill_x_synthetic(hspmv_matrix *A, int64_t global_offset) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  const int grid = grid_for(A->n, 256, A->dev.sms * 16);
  if (A->dtype == HSPMV_F64)
    hspmv::vector_x_kernel<double><<<grid, 256, 0, A->stream>>>(A->n, global_offset, (double *)A->d_x);
  else
    hspmv::vector_x_kernel<float><<<grid, 256, 0, A->stream>>>(A->n, global_offset, (float *)A->d_x);
  LAUNCHED();
  CU(cudaStreamSynchronize(A->stream));
  return HSPMV_OK;
}

int hspmv_fill_x_segments(void *d_x, int dtype, int n_seg, const int64_t *lo, const int64_t *hi, int64_t global_offset,
                          double scale, void *stream) {
  if (!d_x || n_seg < 0 || n_seg > 8 || (n_seg && (!lo || !hi)) || (dtype != HSPMV_F32 && dtype != HSPMV_F64))
    return fail(HSPMV_EINVAL, "bad arguments");
  hspmv::FillSegs segs;
  segs.n = n_seg;
  int64_t longest = 0;
  for (int i = 0; i < n_seg; ++i) {
    if (lo[i] < 0 || hi[i] < lo[i]) return fail(HSPMV_EINVAL, "bad segment");
    segs.lo[i] = lo[i];
    segs.hi[i] = hi[i];
    longest = std::max(longest, hi[i] - lo[i]);
  }
  if (!longest) return HSPMV_OK;
  const int grid = grid_for(longest, 256, 148 * 16);
  if (dtype == HSPMV_F64)
    hspmv::vector_x_segments_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(segs, global_offset, scale, (double *)d_x);
  else
    hspmv::vector_x_segments_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(segs, global_offset, scale, (float *)d_x);
  LAUNCHED();
  CU(cudaGetLastError());
  return HSPMV_OK;
}

// ---- hot path ---------------------------------------------------------------------------
int hspmv_set_option(hspmv_matrix *A, int option, int64_t value) {
  if (!A || option < 1 || option > 31) return fail(HSPMV_EINVAL, "bad option");
  if (option == HSPMV_OPT_BLOCK_NNZ && value != 0 && (value < 256 || value > 16384))
    return fail(HSPMV_EINVAL, "HSPMV_OPT_BLOCK_NNZ must be 0 or within [256, 16384]");
  if (option == HSPMV_OPT_VEC && value != 0 && (value < 1 || value > 32 || (value & (value - 1))))
    return fail(HSPMV_EINVAL, "HSPMV_OPT_VEC must be 0 or a power of two <= 32");
  A->opt[option] = value;
  A->graph_iters = 0; // a captured graph froze the old option
  return HSPMV_OK;
}

int hspmv_spmv(hspmv_matrix *A, int algo, const void *d_x, void *d_y, void *stream) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  return do_spmv(A, algo, d_x, d_y, (cudaStream_t)stream);
}

int hspmv_spmv_range(hspmv_matrix *A, int algo, int64_t row_lo, int64_t row_hi, const void *d_x, void *d_y,
                     void *stream) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  return do_spmv(A, algo, d_x, d_y, (cudaStream_t)stream, row_lo, row_hi);
}

int hspmv_spmv_halo(hspmv_matrix *A, int64_t interior_lo, int64_t interior_hi, const void *d_x, void *d_y,
                    const int *wait_flags, int n_wait, int wait_step, int *timed_out, void *stream) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  if (A->k != 1) return fail(HSPMV_EUNSUPPORTED, "hspmv_spmv_halo works on plain-CSR shards");
  if (interior_lo < 0 || interior_lo > interior_hi || interior_hi > A->m || n_wait < 0 || n_wait > 32 ||
      (n_wait && (!wait_flags || !timed_out)))
    return fail(HSPMV_EINVAL, "bad arguments");
  HaloWait h;
  // interior first, then the rows that need the halo; empty ranges are skipped by the tiling
  h.n_ranges = 3;
  h.lo[0] = interior_lo; h.hi[0] = interior_hi;
  h.lo[1] = 0;           h.hi[1] = interior_lo;
  h.lo[2] = interior_hi; h.hi[2] = A->m;
  h.flags = wait_flags;
  h.n_wait = n_wait;
  h.step = wait_step;
  h.timed_out = timed_out;
  A->last_stream = (cudaStream_t)stream;
  A->last_stream_valid = true;
  int rc = ensure_stats(A);
  if (rc) return rc;
  if (A->dtype == HSPMV_F64)
    rc = run_csrk<double>(A, true, (const double *)(d_x ? d_x : A->d_x), (double *)(d_y ? d_y : A->d_y), (cudaStream_t)stream, 0, A->m, &h);
  else
    rc = run_csrk<float>(A, true, (const float *)(d_x ? d_x : A->d_x), (float *)(d_y ? d_y : A->d_y), (cudaStream_t)stream, 0, A->m, &h);
  if (rc) return rc;
  CU(cudaGetLastError());
  return HSPMV_OK;
}

// ---- stream-ordered flag operations executed by the front end (no SM): the halo protocol ----
namespace {
typedef int (*StreamValueFn)(void *stream, unsigned long long addr, unsigned int value, unsigned int flags);
StreamValueFn g_wait_value = nullptr, g_write_value = nullptr;
int load_stream_memops() {
  if (g_wait_value && g_write_value) return HSPMV_OK;
  void *w = nullptr, *r = nullptr;
  cudaDriverEntryPointQueryResult q1, q2;
  CU(cudaGetDriverEntryPoint("cuStreamWaitValue32", &w, cudaEnableDefault, &q1));
  CU(cudaGetDriverEntryPoint("cuStreamWriteValue32", &r, cudaEnableDefault, &q2));
  if (!w || !r || q1 != cudaDriverEntryPointSuccess || q2 != cudaDriverEntryPointSuccess)
    return fail(HSPMV_EUNSUPPORTED, "driver has no stream memory operations");
  g_wait_value = (StreamValueFn)w;
  g_write_value = (StreamValueFn)r;
  return HSPMV_OK;
}
} // namespace

int hspmv_stream_wait_ge(void *stream, const int *addr, int value) {
  int rc = load_stream_memops();
  if (rc) return rc;
  const int e = g_wait_value(stream, (unsigned long long)(uintptr_t)addr, (unsigned int)value, 0x1 /* CU_STREAM_WAIT_VALUE_GEQ */);
  if (e) return fail(HSPMV_ECUDA, "cuStreamWaitValue32 failed with CUresult " + std::to_string(e));
  return HSPMV_OK;
}
int hspmv_stream_write(void *stream, int *addr, int value) {
  int rc = load_stream_memops();
  if (rc) return rc;
  // default flags: the write is ordered after everything that precedes it in the stream (the
  // peer copies / the kernel).  (CU_STREAM_WRITE_VALUE_NO_MEMORY_BARRIER is rejected with
  // CUDA_ERROR_INVALID_VALUE by driver 580 through this entry point, and it would also allow
  // the flag to pass the payload.)
  const int e = g_write_value(stream, (unsigned long long)(uintptr_t)addr, (unsigned int)value, 0);
  if (e) return fail(HSPMV_ECUDA, "cuStreamWriteValue32 failed with CUresult " + std::to_string(e));
  return HSPMV_OK;
}
int hspmv_memcpy_async(void *dst, const void *src, int64_t bytes, void *stream) {
  CU(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return HSPMV_OK;
}

int hspmv_spmv_host_timed(hspmv_matrix *A, int algo, const void *x_host, void *y_host, int iters,
                          float *ms_total) {
  if (!A || !x_host || !y_host || iters < 1 || !ms_total) return fail(HSPMV_EINVAL, "bad arguments");
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, A->stream));
  for (int i = 0; i < iters; ++i) {
    int rc = hspmv_spmv_host(A, algo, x_host, y_host);
    if (rc) return rc;
  }
  CU(cudaEventRecord(e1, A->stream));
  CU(cudaEventSynchronize(e1));
  CU(cudaEventElapsedTime(ms_total, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return HSPMV_OK;
}

int hspmv_spmv_timed(hspmv_matrix *A, int algo, int warmup, int iters, int flush_l2, float *ms_out) {
  if (!A || iters < 1 || !ms_out) return fail(HSPMV_EINVAL, "bad arguments");
  if (flush_l2 && !A->d_flush) {
    A->flush_bytes = (size_t)A->dev.l2 * 2;
    CU(cudaMalloc(&A->d_flush, A->flush_bytes));
  }
  if (flush_l2 == 2 && !A->d_flush_rd) { // second buffer, only read: evicts the first one's dirty lines
    CU(cudaMalloc(&A->d_flush_rd, A->flush_bytes + 64));
    CU(cudaMemset(A->d_flush_rd, 0, A->flush_bytes + 64));
  }
  for (int i = 0; i < warmup; ++i) {
    int rc = do_spmv(A, algo, nullptr, nullptr, A->stream);
    if (rc) return rc;
  }
  std::vector<cudaEvent_t> ev((size_t)iters * 2);
  for (auto &e : ev) CU(cudaEventCreate(&e));
  for (int i = 0; i < iters; ++i) {
    if (flush_l2) CU(cudaMemsetAsync(A->d_flush, i & 0xff, A->flush_bytes, A->stream));
    if (flush_l2 == 2) {
      hspmv::flush_read_kernel<<<A->dev.sms * 8, 256, 0, A->stream>>>((const int4 *)A->d_flush_rd, (int64_t)(A->flush_bytes / 16),
                                                                       (int *)((char *)A->d_flush_rd + A->flush_bytes));
      LAUNCHED();
    }
    CU(cudaEventRecord(ev[(size_t)2 * i], A->stream));
    int rc = do_spmv(A, algo, nullptr, nullptr, A->stream);
    if (rc) return rc;
    CU(cudaEventRecord(ev[(size_t)2 * i + 1], A->stream));
  }
  CU(cudaStreamSynchronize(A->stream));
  for (int i = 0; i < iters; ++i) CU(cudaEventElapsedTime(&ms_out[i], ev[(size_t)2 * i], ev[(size_t)2 * i + 1]));
  for (auto &e : ev) cudaEventDestroy(e);
  return HSPMV_OK;
}

// Solver-style loops on matrices whose product takes tens of microseconds: `iters` products on the handle's own x / y,
// captured ONCE as a CUDA graph (per handle, algo and iters; kernel options are frozen at capture) and replayed with one
// host call, so neither the per-launch host cost nor the stream's launch-to-launch gap is paid `iters` times.
int hspmv_spmv_graph(hspmv_matrix *A, int algo, int iters, void *stream) {
  if (!A || iters < 1) return fail(HSPMV_EINVAL, "bad arguments");
  cudaStream_t st = stream ? (cudaStream_t)stream : A->stream;
  if (!A->graph_exec || A->graph_algo != algo || A->graph_iters != iters) {
    if (A->graph_exec) cudaGraphExecDestroy(A->graph_exec);
    A->graph_exec = nullptr;
    int rc = do_spmv(A, algo, nullptr, nullptr, A->stream); // uncaptured first: lazy tables, attributes, module load
    if (rc) return rc;
    CU(cudaStreamSynchronize(A->stream));
    const int64_t before = g_launches.load();
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(A->stream, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < iters && !rc; ++i) rc = do_spmv(A, algo, nullptr, nullptr, A->stream);
    cudaError_t e = cudaStreamEndCapture(A->stream, &graph);
    A->graph_launches = g_launches.load() - before;
    g_launches.fetch_sub(A->graph_launches, std::memory_order_relaxed); // captured, not launched
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (e != cudaSuccess) return fail(HSPMV_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&A->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(HSPMV_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    A->graph_algo = algo;
    A->graph_iters = iters;
  }
  CU(cudaGraphLaunch(A->graph_exec, st));
  g_launches.fetch_add(A->graph_launches, std::memory_order_relaxed);
  A->last_stream = st;
  A->last_stream_valid = true;
  return HSPMV_OK;
}

// ms per product of `reps` replays of the `iters`-product graph
int hspmv_spmv_graph_timed(hspmv_matrix *A, int algo, int iters, int reps, float *ms_per_product) {
  if (!A || iters < 1 || reps < 1 || !ms_per_product) return fail(HSPMV_EINVAL, "bad arguments");
  int rc = hspmv_spmv_graph(A, algo, iters, A->stream); // capture + one warm replay
  if (rc) return rc;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, A->stream));
  for (int r = 0; r < reps; ++r)
    if ((rc = hspmv_spmv_graph(A, algo, iters, A->stream))) return rc;
  CU(cudaEventRecord(e1, A->stream));
  CU(cudaStreamSynchronize(A->stream));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_product = ms / (float)(iters * reps);
  return HSPMV_OK;
}

int hspmv_spmv_timed_batch(hspmv_matrix *A, int algo, int warmup, int iters, float *ms_per_launch) {
  if (!A || iters < 1 || !ms_per_launch) return fail(HSPMV_EINVAL, "bad arguments");
  for (int i = 0; i < warmup; ++i) {
    int rc = do_spmv(A, algo, nullptr, nullptr, A->stream);
    if (rc) return rc;
  }
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, A->stream));
  for (int i = 0; i < iters; ++i) {
    int rc = do_spmv(A, algo, nullptr, nullptr, A->stream);
    if (rc) return rc;
  }
  CU(cudaEventRecord(e1, A->stream));
  CU(cudaStreamSynchronize(A->stream));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_launch = ms / (float)iters;
  return HSPMV_OK;
}

namespace {

// Resolve AUTO the way dispatch() does; the pipelined host path needs a kernel that takes row ranges.
int resolve_algo(hspmv_matrix *A, int algo) {
  if (algo != HSPMV_ALGO_AUTO) return algo;
  if (A->k >= 2) return HSPMV_ALGO_CSRK;
  if (ensure_stats(A)) return HSPMV_ALGO_CSR_VECTOR;
  const double mean = A->m ? (double)A->nnz / (double)A->m : 0.0;
  return (A->stats[8] <= 8 * (mean + 1) + 64) ? HSPMV_ALGO_CSR_STREAM : HSPMV_ALGO_CSR_VECTOR;
}

int build_pipe(hspmv_matrix *A) {
  hspmv_matrix::Pipe &P = A->pipe;
  P.tried = true;
  const size_t vs = A->vsize();
  if (A->d_perm || A->k != 1 || (size_t)A->m * vs < ((size_t)32 << 20)) return HSPMV_OK; // small or permuted: plain path
  // equal row chunks, and x pieces cut where the chunks' column ranges end (partition.cpp: pipe_chunks, pipe_pieces --
  // host rules, property-tested without a device through hspmv_pipe_plan_host)
  hspmv::pipe_chunks(A->m, (int64_t)vs, P.row_lo);
  const int C = (int)P.row_lo.size() - 1;
  P.chunks = C;
  // highest column each row chunk references -> the last x piece it has to wait for
  std::vector<int> e((size_t)C + 1);
  for (int c = 0; c <= C; ++c)
    CU(cudaMemcpy(&e[(size_t)c], A->d_rp + P.row_lo[(size_t)c], 4, cudaMemcpyDeviceToHost));
  int *d = nullptr;
  CU(cudaMalloc(&d, (size_t)C * 8));
  std::vector<int> init((size_t)C * 2);
  for (int c = 0; c < C; ++c) {
    init[(size_t)2 * c] = INT32_MAX;
    init[(size_t)2 * c + 1] = -1;
  }
  CU(cudaMemcpy(d, init.data(), (size_t)C * 8, cudaMemcpyHostToDevice));
  for (int c = 0; c < C; ++c) {
    const int64_t cnt = (int64_t)e[(size_t)c + 1] - e[(size_t)c];
    if (cnt <= 0) continue;
    hspmv::col_minmax_kernel<<<grid_for(cnt, 256, A->dev.sms * 8), 256, 0, A->stream>>>(cnt, A->d_ci + e[(size_t)c], d + 2 * c, d + 2 * c + 1);
    LAUNCHED();
  }
  CU(cudaStreamSynchronize(A->stream));
  CU(cudaMemcpy(init.data(), d, (size_t)C * 8, cudaMemcpyDeviceToHost));
  cudaFree(d);
  std::vector<int64_t> cmax((size_t)C);
  for (int c = 0; c < C; ++c) cmax[(size_t)c] = init[(size_t)2 * c + 1];
  hspmv::pipe_pieces(A->n, cmax, P.x_lo, P.need);
  CU(cudaStreamCreateWithFlags(&P.s_h2d, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&P.s_d2h, cudaStreamNonBlocking));
  P.ev_x.resize((size_t)C);
  P.ev_k.resize((size_t)C);
  for (int c = 0; c < C; ++c) {
    CU(cudaEventCreateWithFlags(&P.ev_x[(size_t)c], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&P.ev_k[(size_t)c], cudaEventDisableTiming));
  }
  P.ok = true;
  return HSPMV_OK;
}

int spmv_host_pipelined(hspmv_matrix *A, int algo, const void *x_host, void *y_host) {
  hspmv_matrix::Pipe &P = A->pipe;
  const size_t vs = A->vsize();
  const int C = P.chunks;
  for (int p = 0; p < C; ++p) {
    const int64_t a = P.x_lo[(size_t)p], b = P.x_lo[(size_t)p + 1];
    if (b > a)
      CU(cudaMemcpyAsync((char *)A->d_x + (size_t)a * vs, (const char *)x_host + (size_t)a * vs, (size_t)(b - a) * vs,
                         cudaMemcpyHostToDevice, P.s_h2d));
    CU(cudaEventRecord(P.ev_x[(size_t)p], P.s_h2d));
  }
  int waited = -1;
  for (int c = 0; c < C; ++c) {
    const int64_t a = P.row_lo[(size_t)c], b = P.row_lo[(size_t)c + 1];
    if (P.need[(size_t)c] > waited) { // pieces arrive in order: waiting for the last one needed is enough
      waited = P.need[(size_t)c];
      CU(cudaStreamWaitEvent(A->stream, P.ev_x[(size_t)waited], 0));
    }
    if (b > a) {
      int rc = do_spmv(A, algo, nullptr, nullptr, A->stream, a, b);
      if (rc) return rc;
    }
    CU(cudaEventRecord(P.ev_k[(size_t)c], A->stream));
    CU(cudaStreamWaitEvent(P.s_d2h, P.ev_k[(size_t)c], 0));
    if (b > a)
      CU(cudaMemcpyAsync((char *)y_host + (size_t)a * vs, (const char *)A->d_y + (size_t)a * vs, (size_t)(b - a) * vs,
                         cudaMemcpyDeviceToHost, P.s_d2h));
  }
  CU(cudaStreamSynchronize(P.s_d2h));
  CU(cudaStreamSynchronize(A->stream));
  return HSPMV_OK;
}

} // namespace

int hspmv_autotune(hspmv_matrix *A, int algo, int effort, float *best_ms) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  const int resolved = resolve_algo(A, algo);
  float ms = 0;
  int rc;
  if (resolved != HSPMV_ALGO_CSR_STREAM && resolved != HSPMV_ALGO_CSRK) { // nothing to choose between
    rc = hspmv_spmv_timed_batch(A, resolved, 2, 5, &ms);
    if (!rc && best_ms) *best_ms = ms;
    return rc;
  }
  if ((rc = ensure_stats(A))) return rc;
  const bool stream = resolved == HSPMV_ALGO_CSR_STREAM;
  const double mean = A->m ? (double)A->nnz / (double)A->m : 1.0;
  struct Cand {
    int64_t vec, chunk, tail, rpt;
  };
  std::vector<Cand> cands;
  cands.push_back({0, 0, 0, 0}); // the heuristic's own choice
  int v0 = 1;
  while (v0 < 32 && mean / v0 > 14.0) v0 <<= 1;
  for (int vec = v0 > 1 ? v0 / 2 : 1; vec <= 32 && vec <= v0 * 2; vec <<= 1) {
    const int per = (int)(mean / vec);
    if (per >= 3 && per <= 13) cands.push_back({vec, per, 0, 0});
    if (effort >= 2 && per - 1 >= 3 && per - 1 <= 13) cands.push_back({vec, per - 1, 0, 0});
    for (int tail : {8, 12}) {
      cands.push_back({vec, -1, tail, stream ? 32 / vec : 0});
      if (effort >= 2 && stream) cands.push_back({vec, -1, tail, 2 * (32 / vec)});
    }
  }
  const int keys[4] = {HSPMV_OPT_VEC, HSPMV_OPT_CHUNK, HSPMV_OPT_TAIL, HSPMV_OPT_ROWS_PER_TILE};
  Cand best = {A->opt[keys[0]], A->opt[keys[1]], A->opt[keys[2]], A->opt[keys[3]]};
  float best_t = -1;
  for (const Cand &c : cands) {
    const int64_t vals[4] = {c.vec, c.chunk, c.tail, c.rpt};
    for (int i = 0; i < 4; ++i) A->opt[keys[i]] = vals[i];
    if (hspmv_spmv_timed_batch(A, resolved, 2, 5, &ms) != HSPMV_OK) { // e.g. no compiled variant
      cudaGetLastError();
      continue;
    }
    if (best_t < 0 || ms < best_t) {
      best_t = ms;
      best = c;
    }
  }
  const int64_t vals[4] = {best.vec, best.chunk, best.tail, best.rpt};
  for (int i = 0; i < 4; ++i) A->opt[keys[i]] = vals[i];
  if (best_t < 0) return fail(HSPMV_ECUDA, "autotune: no kernel shape ran");
  rc = hspmv_spmv_timed_batch(A, resolved, 1, 3, &ms); // leaves last_kernel describing the winner
  if (!rc && best_ms) *best_ms = ms < best_t ? ms : best_t;
  return rc;
}

int hspmv_spmv_host(hspmv_matrix *A, int algo, const void *x_host, void *y_host) {
  if (!A || !x_host || !y_host) return fail(HSPMV_EINVAL, "null");
  const int resolved = resolve_algo(A, algo);
  // (the pipelined form launches row CHUNKS; only the tile engine takes them at full speed -- a binned matrix would run its
  // chunks through the contiguous vector kernel, which is what the bins exist to avoid)
  if (resolved == HSPMV_ALGO_CSR_STREAM && !A->opt[HSPMV_OPT_NO_PIPELINE]) {
    if (!A->pipe.tried) {
      int rc0 = build_pipe(A);
      if (rc0) return rc0;
    }
    if (A->pipe.ok) return spmv_host_pipelined(A, resolved, x_host, y_host);
  }
  int rc = hspmv_set_x(A, x_host);
  if (rc) return rc;
  rc = do_spmv(A, algo, nullptr, nullptr, A->stream);
  if (rc) return rc;
  return hspmv_get_y(A, y_host, 1);
}

const char *hspmv_last_kernel(const hspmv_matrix *A) { return A ? A->last_kernel.c_str() : ""; }

// ---- spmv.h compatibility ------------------------------------------------------------------
// omp_spmv(m, n, nnz, row_start, col_idx, val, x, y) carries no handle, so the uploaded matrix is
// cached.  The key is the arrays' identity AND a content fingerprint (all of row_start, 4096
// evenly spaced samples of col_idx and val, both ends included), checked on every call: a caller
// that refills the arrays in place, or frees them and allocates another matrix at the same
// addresses, gets a fresh upload instead of the stale device matrix.  At most COMPAT_CACHE_MAX
// matrices stay resident (least recently used goes first).
namespace {
struct CompatKey {
  const void *rs, *ci, *val;
  int m, nnz, dtype;
  bool operator<(const CompatKey &o) const {
    return std::tie(rs, ci, val, m, nnz, dtype) < std::tie(o.rs, o.ci, o.val, o.m, o.nnz, o.dtype);
  }
};
struct CompatEntry {
  hspmv_matrix *A;
  uint64_t print, used;
};
std::map<CompatKey, CompatEntry> g_cache;
std::mutex g_cache_mu;
uint64_t g_cache_clock = 0;
constexpr size_t COMPAT_CACHE_MAX = 4;

uint64_t fnv(uint64_t h, const void *p, size_t bytes) {
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < bytes; ++i) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}
uint64_t compat_fingerprint(int dtype, int m, int nnz, const int *rs, const int *ci, const void *val) {
  const size_t vs = dtype == HSPMV_F64 ? 8 : 4;
  uint64_t h = 1469598103934665603ull;
  // row_start in 64-bit words (fast), the odd tail by bytes
  const size_t words = ((size_t)m + 1) / 2;
  const uint64_t *w = (const uint64_t *)rs;
  for (size_t i = 0; i < words; ++i) h = (h ^ w[i]) * 1099511628211ull;
  h = fnv(h, rs + words * 2, ((size_t)m + 1 - words * 2) * 4);
  if (nnz > 0) {
    const int64_t samples = nnz < 4096 ? nnz : 4096;
    for (int64_t s = 0; s < samples; ++s) {
      const int64_t j = samples > 1 ? (int64_t)(nnz - 1) * s / (samples - 1) : 0;
      h = fnv(h, ci + j, 4);
      h = fnv(h, (const char *)val + (size_t)j * vs, vs);
    }
  }
  return h;
}

int compat_spmv(int dtype, int m, int n, int nnz, const int *rs, const int *ci, const void *val,
                const void *x, void *y) {
  if (!rs || m < 0 || nnz < 0 || (nnz && (!ci || !val))) return fail(HSPMV_EINVAL, "null CSR array");
  std::lock_guard<std::mutex> lock(g_cache_mu);
  const CompatKey key{rs, ci, val, m, nnz, dtype};
  const uint64_t print = compat_fingerprint(dtype, m, nnz, rs, ci, val);
  hspmv_matrix *A = nullptr;
  auto it = g_cache.find(key);
  if (it != g_cache.end() && it->second.print != print) { // same addresses, different matrix
    hspmv_destroy(it->second.A);
    g_cache.erase(it);
    it = g_cache.end();
  }
  if (it == g_cache.end()) {
    while (g_cache.size() >= COMPAT_CACHE_MAX) {
      auto old = g_cache.begin();
      for (auto j = g_cache.begin(); j != g_cache.end(); ++j)
        if (j->second.used < old->second.used) old = j;
      hspmv_destroy(old->second.A);
      g_cache.erase(old);
    }
    // the reference's omp_spmv indexes x with column ids < m for its square inputs; n is unused there
    int rc = hspmv_create_csr(&A, m, n > 0 ? n : m, nnz, rs, ci, val, dtype, 0, 0);
    if (rc) return rc;
    g_cache[key] = CompatEntry{A, print, ++g_cache_clock};
  } else {
    A = it->second.A;
    it->second.used = ++g_cache_clock;
  }
  return hspmv_spmv_host(A, HSPMV_ALGO_AUTO, x, y);
}
} // namespace

int hspmv_omp_spmv_f64(int m, int n, int nnz, const int *row_start, const int *col_idx, const double *val,
                       const double *x, double *y) {
  return compat_spmv(HSPMV_F64, m, n, nnz, row_start, col_idx, val, x, y);
}
int hspmv_omp_spmv_f32(int m, int n, int nnz, const int *row_start, const int *col_idx, const float *val,
                       const float *x, float *y) {
  return compat_spmv(HSPMV_F32, m, n, nnz, row_start, col_idx, val, x, y);
}
int hspmv_omp_spmv_cache_size(void) {
  std::lock_guard<std::mutex> lock(g_cache_mu);
  return (int)g_cache.size();
}
void hspmv_omp_spmv_cache_clear(void) {
  std::lock_guard<std::mutex> lock(g_cache_mu);
  for (auto &kv : g_cache) hspmv_destroy(kv.second.A);
  g_cache.clear();
}

// ---- multi-GPU index work ----------------------------------------------------------------
int hspmv_partition_rows(int64_t m, const int32_t *row_ptr, int n_parts, const uint32_t *align_ptr,
                         int64_t n_align, int64_t *row_lo) {
  int rc = hspmv::partition_rows(m, row_ptr, n_parts, align_ptr, n_align, row_lo);
  return rc ? fail(rc, "partition_rows: bad arguments") : rc;
}
int hspmv_block_col_range(const int32_t *row_ptr, const int32_t *col_idx, int64_t row_lo, int64_t row_hi,
                          int64_t *col_min, int64_t *col_max) {
  int rc = hspmv::block_col_range(row_ptr, col_idx, row_lo, row_hi, col_min, col_max);
  return rc ? fail(rc, "block_col_range: bad arguments") : rc;
}
int hspmv_slice_rows(const hspmv_matrix *A, int64_t row_lo, int64_t row_hi, int64_t col_lo, int64_t n_cols,
                     hspmv_matrix **out) {
  if (!A || !out) return fail(HSPMV_EINVAL, "null");
  if (row_lo < 0 || row_lo > row_hi || row_hi > A->m) return fail(HSPMV_EINVAL, "bad row range");
  if (n_cols <= 0) {
    col_lo = 0;
    n_cols = A->n;
  }
  if (col_lo < 0 || col_lo + n_cols > A->n) return fail(HSPMV_EINVAL, "bad column window");
  // CSR-k: the cut must not split a super-row (k = 2) / super-super-row (k = 3)
  int64_t s0 = 0, s1 = 0, t0 = 0, t1 = 0;
  if (A->k >= 2) {
    auto find = [](const uvec &v, int64_t x) -> int64_t {
      auto it = std::lower_bound(v.begin(), v.end(), (u32)x);
      return (it != v.end() && *it == (u32)x) ? (int64_t)(it - v.begin()) : -1;
    };
    s0 = find(A->h_sr, row_lo);
    s1 = find(A->h_sr, row_hi);
    if (s0 < 0 || s1 < 0) return fail(HSPMV_EINVAL, "slice bounds are not super-row boundaries");
    if (A->k == 3) {
      t0 = find(A->h_ssr, s0);
      t1 = find(A->h_ssr, s1);
      if (t0 < 0 || t1 < 0) return fail(HSPMV_EINVAL, "slice bounds are not super-super-row boundaries");
    }
  }
  int e[2] = {0, 0};
  CU(cudaMemcpy(&e[0], A->d_rp + row_lo, 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&e[1], A->d_rp + row_hi, 4, cudaMemcpyDeviceToHost));
  hspmv_matrix *B = nullptr;
  int rc = new_handle(&B, A->dtype, row_hi - row_lo, n_cols, (int64_t)e[1] - e[0]);
  if (rc) return rc;
  if ((rc = alloc_matrix_arrays(B)) || (rc = alloc_vectors(B))) {
    hspmv_destroy(B);
    return rc;
  }
  int *d_bad = nullptr;
  cudaError_t ce = cudaMalloc(&d_bad, 4);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(d_bad, 0, 4, B->stream);
  if (ce == cudaSuccess) {
    hspmv::rebase_kernel<<<grid_for(B->m + 1, 256, B->dev.sms * 16), 256, 0, B->stream>>>(B->m + 1, A->d_rp + row_lo, e[0], e[1] - e[0] + 1,
                                                                                     B->d_rp, d_bad);
    LAUNCHED();
    if (B->nnz) {
      hspmv::rebase_kernel<<<grid_for(B->nnz, 256, B->dev.sms * 16), 256, 0, B->stream>>>(B->nnz, A->d_ci + e[0], (int)col_lo, (int)n_cols,
                                                                                     B->d_ci, d_bad);
      LAUNCHED();
      ce = cudaMemcpyAsync(B->d_val, (const char *)A->d_val + (size_t)e[0] * A->vsize(), (size_t)B->nnz * A->vsize(),
                           cudaMemcpyDeviceToDevice, B->stream);
    }
  }
  int bad = 0;
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, B->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(B->stream);
  cudaFree(d_bad);
  if (ce != cudaSuccess || bad) {
    hspmv_destroy(B);
    if (bad) return fail(HSPMV_EINVAL, "the row block references a column outside the window");
    return fail(HSPMV_ECUDA, std::string("slice_rows: ") + cudaGetErrorString(ce));
  }
  B->h_rp.resize((size_t)B->m + 1);
  ce = cudaMemcpy(B->h_rp.data(), B->d_rp, (size_t)(B->m + 1) * 4, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) {
    hspmv_destroy(B);
    return fail(HSPMV_ECUDA, std::string("slice_rows: ") + cudaGetErrorString(ce));
  }
  hspmv::row_stats(B->m, B->h_rp.data(), B->stats);
  B->stats_ready = true;
  if (A->k >= 2) {
    B->k = A->k;
    B->h_sr.resize((size_t)(s1 - s0) + 1);
    for (int64_t i = s0; i <= s1; ++i) B->h_sr[(size_t)(i - s0)] = A->h_sr[(size_t)i] - (u32)row_lo;
    B->n_sr = s1 - s0;
    if (A->k == 3) {
      B->h_ssr.resize((size_t)(t1 - t0) + 1);
      for (int64_t i = t0; i <= t1; ++i) B->h_ssr[(size_t)(i - t0)] = A->h_ssr[(size_t)i] - (u32)s0;
      B->n_ssr = t1 - t0;
    }
    if (B->n_sr > 0 && (rc = upload_csrk_maps(B))) {
      hspmv_destroy(B);
      return rc;
    }
    if (B->n_sr == 0) B->k = 1; // an empty slice has no super-rows to walk
  }
  *out = B;
  return HSPMV_OK;
}

int hspmv_col_range_device(const hspmv_matrix *A, int64_t *col_min, int64_t *col_max) {
  if (!A || !col_min || !col_max) return fail(HSPMV_EINVAL, "null");
  int *d = nullptr;
  CU(cudaMalloc(&d, 8));
  int init[2] = {INT32_MAX, -1};
  CU(cudaMemcpy(d, init, 8, cudaMemcpyHostToDevice));
  if (A->nnz) {
    hspmv::col_minmax_kernel<<<grid_for(A->nnz, 256, A->dev.sms * 16), 256, 0, A->stream>>>(A->nnz, A->d_ci, d, d + 1);
    LAUNCHED();
  }
  int res[2];
  CU(cudaStreamSynchronize(A->stream));
  CU(cudaMemcpy(res, d, 8, cudaMemcpyDeviceToHost));
  cudaFree(d);
  *col_min = A->nnz ? res[0] : 0;
  *col_max = A->nnz ? res[1] : -1;
  return HSPMV_OK;
}

int hspmv_interior_rows(const hspmv_matrix *A, int64_t own_col_lo, int64_t own_col_hi, int64_t *row_a,
                        int64_t *row_b) {
  if (!A || !row_a || !row_b) return fail(HSPMV_EINVAL, "null");
  int *d = nullptr;
  CU(cudaMalloc(&d, 8));
  int init[2] = {0, (int)A->m};
  CU(cudaMemcpy(d, init, 8, cudaMemcpyHostToDevice));
  if (A->m) {
    hspmv::interior_rows_kernel<<<grid_for(A->m * 32, 256, A->dev.sms * 16), 256, 0, A->stream>>>(
        A->m, A->d_rp, A->d_ci, (int)own_col_lo, (int)own_col_hi, d, d + 1);
    LAUNCHED();
  }
  CU(cudaStreamSynchronize(A->stream));
  int res[2];
  CU(cudaMemcpy(res, d, 8, cudaMemcpyDeviceToHost));
  cudaFree(d);
  *row_a = res[0];
  *row_b = res[1] < res[0] ? res[0] : res[1];
  return HSPMV_OK;
}

// ---- NVLink peer exchange (CUDA IPC) -------------------------------------------------------
int hspmv_peer_export(void *dptr, void *handle64) {
  if (!dptr || !handle64) return fail(HSPMV_EINVAL, "null");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, dptr));
  return HSPMV_OK;
}
int hspmv_peer_open(const void *handle64, void **dptr) {
  if (!handle64 || !dptr) return fail(HSPMV_EINVAL, "null");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, sizeof(h));
  CU(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return HSPMV_OK;
}
int hspmv_peer_close(void *dptr) {
  if (dptr) CU(cudaIpcCloseMemHandle(dptr));
  return HSPMV_OK;
}
int hspmv_peer_alloc(int64_t bytes, void **dptr) {
  if (!dptr || bytes <= 0) return fail(HSPMV_EINVAL, "bad arguments");
  CU(cudaMalloc(dptr, (size_t)bytes));
  CU(cudaMemset(*dptr, 0, (size_t)bytes));
  return HSPMV_OK;
}
int hspmv_peer_free(void *dptr) {
  if (dptr) CU(cudaFree(dptr));
  return HSPMV_OK;
}

int64_t hspmv_row_len_percentile_host(int64_t m, const int32_t *row_ptr, int pct) {
  if (!row_ptr || m < 0 || pct < 0 || pct > 100) return -1;
  return hspmv::row_len_percentile(m, row_ptr, pct);
}

// tile table of the CSR_MERGE path, host restatement (for the bit-exactness tests)
int hspmv_merge_tile_rows_host(int64_t m, const int32_t *row_ptr, int64_t tile, int32_t *tile_row) {
  if (!row_ptr || !tile_row || tile <= 0) return fail(HSPMV_EINVAL, "bad arguments");
  std::vector<int32_t> t;
  hspmv::tile_rows(m, row_ptr, tile, t);
  std::memcpy(tile_row, t.data(), t.size() * 4);
  return HSPMV_OK;
}
// device tile table of a handle (after a CSR_MERGE launch), copied back for diffing
int hspmv_merge_tile_rows_device(const hspmv_matrix *A, int64_t *tile, int64_t *n_tiles, int32_t *tile_row) {
  if (!A || !A->d_tile_row) return fail(HSPMV_ESTATE, "no CSR_MERGE launch on this handle yet");
  if (tile) *tile = A->merge_tile;
  if (n_tiles) *n_tiles = A->merge_tiles;
  if (tile_row) CU(cudaMemcpy(tile_row, A->d_tile_row, (size_t)(A->merge_tiles + 1) * 4, cudaMemcpyDeviceToHost));
  return HSPMV_OK;
}
// row blocks of the CSR_BLOCK path, host rule: blk_row has room for m + 2 entries; *n_blocks out
int hspmv_block_rows_host(int64_t m, const int32_t *row_ptr, int64_t cap, int64_t max_rows, int32_t *blk_row, int64_t *n_blocks) {
  if (!row_ptr || !blk_row || !n_blocks || cap <= 0 || max_rows <= 0) return fail(HSPMV_EINVAL, "bad arguments");
  std::vector<int32_t> b;
  hspmv::block_rows(m, row_ptr, cap, max_rows, b);
  std::memcpy(blk_row, b.data(), b.size() * 4);
  *n_blocks = (int64_t)b.size() - 1;
  return HSPMV_OK;
}
// Band-k grouping rule, both statements (for the tests): start[] has room for n + 2 entries, grp[n]
int hspmv_bandk_group_rows_host(int64_t n, const uint32_t *row_ptr, uint32_t threshold, int literal, uint32_t *start,
                                int64_t *n_starts, uint32_t *grp) {
  if (!row_ptr || !start || !n_starts || !grp || n < 0) return fail(HSPMV_EINVAL, "bad arguments");
  uvec s, g;
  if (literal)
    hspmv::bandk_group_rows_literal(row_ptr, n, threshold, s, g);
  else
    hspmv::bandk_group_rows_fast(row_ptr, n, threshold, s, g);
  std::memcpy(start, s.data(), s.size() * 4);
  if (n) std::memcpy(grp, g.data(), (size_t)n * 4);
  *n_starts = (int64_t)s.size();
  return HSPMV_OK;
}
// column order of the hub relabel, host rule (for the tests): perm[new] = old, inv[old] = new
int hspmv_hub_order_host(int64_t n, const uint32_t *count, uint32_t *perm, uint32_t *inv) {
  if (n < 0 || (n > 0 && (!count || !perm || !inv))) return fail(HSPMV_EINVAL, "bad arguments");
  uvec p, q;
  hspmv::hub_order(n, count, p, q);
  if (n) std::memcpy(perm, p.data(), (size_t)n * 4);
  if (n) std::memcpy(inv, q.data(), (size_t)n * 4);
  return HSPMV_OK;
}
// plan of the pipelined host call, host rules (for the property tests): row_lo / x_lo have room for 66 entries, need for 65
int hspmv_pipe_plan_host(int64_t m, int64_t n, int vsize, const int64_t *chunk_cmax, int64_t *row_lo, int *n_chunks, int64_t *x_lo,
                         int *need) {
  if (!row_lo || !n_chunks || (vsize != 4 && vsize != 8)) return fail(HSPMV_EINVAL, "bad arguments");
  std::vector<int64_t> rl, xl;
  std::vector<int> nd;
  hspmv::pipe_chunks(m, vsize, rl);
  const int C = (int)rl.size() - 1;
  if (C > 64) return fail(HSPMV_EINVAL, "more than 64 chunks");
  std::memcpy(row_lo, rl.data(), rl.size() * 8);
  *n_chunks = C;
  if (chunk_cmax && x_lo && need) {
    std::vector<int64_t> cm(chunk_cmax, chunk_cmax + C);
    hspmv::pipe_pieces(n, cm, xl, nd);
    std::memcpy(x_lo, xl.data(), xl.size() * 8);
    std::memcpy(need, nd.data(), nd.size() * 4);
  }
  return HSPMV_OK;
}
// row bins of the CSR_VECTOR path, host (for the bit-exactness tests): rows[m], bin_ptr[9]
int hspmv_bin_rows_host(int64_t m, const int32_t *row_ptr, int32_t *rows, int64_t *bin_ptr) {
  if (!row_ptr || !rows || !bin_ptr) return fail(HSPMV_EINVAL, "null");
  std::vector<int32_t> r;
  int64_t bp[hspmv::NUM_BINS + 1];
  hspmv::bin_rows(m, row_ptr, r, bp);
  if (m) std::memcpy(rows, r.data(), (size_t)m * 4);
  std::memcpy(bin_ptr, bp, sizeof(bp));
  return HSPMV_OK;
}

} // extern "C"

// ---- STS: sparse triangular solve on the level-set-ordered matrix ------------------------------
namespace {

// lanes per row of the solve: one round of four entries per lane should cover the lower part of a typical row
int sts_vec(const hspmv_matrix *A) {
  const double mean_lower = A->m ? 0.5 * (double)A->nnz / (double)A->m : 1.0;
  int vec = 1;
  while (vec < 32 && mean_lower / vec > 4.0) vec <<= 1;
  return vec;
}

// the barrier-free solve (sts_flags_kernel): every group resident, rows in increasing order, per-row ready flags
template <typename T>
int launch_sts_flags(hspmv_matrix *A, cudaStream_t st) {
  int vec = sts_vec(A);
  void (*kern)(int, int, const int *, const int *, const int *, const T *, const T *, T *, int *, int *) = nullptr;
  switch (vec) {
  case 1: kern = hspmv::sts_flags_kernel<T, 1>; break;
  case 2: kern = hspmv::sts_flags_kernel<T, 2>; break;
  case 4: kern = hspmv::sts_flags_kernel<T, 4>; break;
  case 8: kern = hspmv::sts_flags_kernel<T, 8>; break;
  case 16: kern = hspmv::sts_flags_kernel<T, 16>; break;
  default: kern = hspmv::sts_flags_kernel<T, 32>; vec = 32; break;
  }
  const int threads = 256;
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
  if (per_sm < 1) return fail(HSPMV_ECUDA, "sts_flags_kernel does not fit on an SM");
  // Resident lane groups: as many as the device takes.  (Fewer groups -- "only as many as can run" -- was measured and is
  // slower at every size: config 3 2.7 ms with all of them, 10.7 ms with 4096; tools/sts_groups_sweep.py.)
  int64_t want_groups = A->opt[HSPMV_OPT_STS_GROUPS];
  if (want_groups <= 0) want_groups = INT32_MAX;
  int64_t grid64 = (want_groups * vec + threads - 1) / threads;
  grid64 = std::max<int64_t>(1, std::min<int64_t>(grid64, (int64_t)A->dev.sms * per_sm));
  int grid = (int)grid64;
  if (++A->sts_epoch == INT32_MAX) { // epochs never repeat within the life of the flag array
    CU(cudaMemsetAsync(A->d_ready, 0, (size_t)(A->m + 1) * 4, st));
    A->sts_epoch = 1;
  }
  int n = (int)A->m, epoch = A->sts_epoch;
  const int *rp = A->d_rp, *diag = A->d_diag, *ci = A->d_ci;
  const T *val = (const T *)A->d_val, *b = (const T *)A->d_sts_b;
  T *x = (T *)A->d_sts_x;
  int *ready = A->d_ready, *timed_out = A->d_ready + A->m;
  void *args[] = {&n, &epoch, &rp, &diag, &ci, &val, &b, &x, &ready, &timed_out};
  CU(cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)grid), dim3((unsigned)threads), args, 0, st)); // (co-residency)
  LAUNCHED();
  char name[128];
  std::snprintf(name, sizeof(name), "sts_flags<%s,vec=%d> rows=%d packs=%d grid=%d threads=%d barrier-free", sizeof(T) == 8 ? "f64" : "f32", vec, n,
                (int)A->n_packs, grid, threads);
  A->last_kernel = name;
  return HSPMV_OK;
}

template <typename T>
int launch_sts_solve(hspmv_matrix *A, cudaStream_t st) {
  int vec = sts_vec(A);
  void (*kern)(int, int, const int *, const int *, const int *, const int *, const T *, const T *, T *) = nullptr;
  switch (vec) {
  case 1: kern = hspmv::sts_solve_kernel<T, 1>; break;
  case 2: kern = hspmv::sts_solve_kernel<T, 2>; break;
  case 4: kern = hspmv::sts_solve_kernel<T, 4>; break;
  case 8: kern = hspmv::sts_solve_kernel<T, 8>; break;
  case 16: kern = hspmv::sts_solve_kernel<T, 16>; break;
  default: kern = hspmv::sts_solve_kernel<T, 32>; vec = 32; break;
  }
  const int threads = 1024;
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
  if (per_sm < 1) return fail(HSPMV_ECUDA, "sts_solve_kernel does not fit on an SM");
  int grid = A->dev.sms; // one CTA of 1024 threads per SM: a grid barrier costs less with fewer CTAs
  int n_packs = (int)A->n_packs, solo = (int)A->solo_packs;
  const int *pack = A->d_pack, *rp = A->d_rp, *diag = A->d_diag, *ci = A->d_ci;
  const T *val = (const T *)A->d_val, *b = (const T *)A->d_sts_b;
  T *x = (T *)A->d_sts_x;
  void *args[] = {&n_packs, &solo, &pack, &rp, &diag, &ci, &val, &b, &x};
  CU(cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)grid), dim3((unsigned)threads), args, 0, st));
  LAUNCHED();
  char name[128];
  std::snprintf(name, sizeof(name), "sts_solve<%s,vec=%d> packs=%d (first %d on one CTA) grid=%d threads=1024 cooperative", sizeof(T) == 8 ? "f64" : "f32",
                vec, n_packs, solo, grid);
  A->last_kernel = name;
  return HSPMV_OK;
}

} // namespace

extern "C" {

int hspmv_sts_levels_host(int64_t n, int64_t nnz, const int32_t *row_ptr, const int32_t *col_idx, uint32_t *perm, int32_t *pack_ptr,
                          int64_t *n_packs) {
  (void)nnz;
  if (!row_ptr || !col_idx || !perm || !pack_ptr || !n_packs) return fail(HSPMV_EINVAL, "null");
  uvec p;
  std::vector<int32_t> pk;
  int rc = hspmv::sts_levels(n, row_ptr, col_idx, p, pk);
  if (rc) return fail(rc, rc == HSPMV_EUNSUPPORTED ? "STS: a row has no stored diagonal, or the level search stalled (unsymmetric structure)" : "STS: bad arguments");
  std::memcpy(perm, p.data(), p.size() * 4);
  std::memcpy(pack_ptr, pk.data(), pk.size() * 4);
  *n_packs = (int64_t)pk.size() - 1;
  return HSPMV_OK;
}

int hspmv_build_sts(hspmv_matrix *A) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  if (A->m != A->n || A->m <= 0) return fail(HSPMV_EINVAL, "STS needs a square matrix");
  if (A->k != 1 || A->sts_ready || A->d_perm) return fail(HSPMV_ESTATE, "matrix is already reordered (CSR-k or STS)");
  const size_t vs = A->vsize();
  if (A->h_rp.empty()) {
    A->h_rp.resize((size_t)A->m + 1);
    hspmv::first_touch(A->h_rp.data(), (size_t)(A->m + 1) * 4);
    CU(cudaMemcpy(A->h_rp.data(), A->d_rp, (size_t)(A->m + 1) * 4, cudaMemcpyDeviceToHost));
  }
  if (A->h_ci.size() != (size_t)A->nnz) {
    A->h_ci.resize((size_t)A->nnz);
    hspmv::first_touch(A->h_ci.data(), (size_t)A->nnz * 4);
    if (A->nnz) CU(cudaMemcpy(A->h_ci.data(), A->d_ci, (size_t)A->nnz * 4, cudaMemcpyDeviceToHost));
  }
  uvec perm;
  std::vector<int32_t> pack;
  int rc = hspmv::sts_levels(A->m, A->h_rp.data(), A->h_ci.data(), perm, pack);
  if (rc) return fail(rc, rc == HSPMV_EUNSUPPORTED ? "STS: a row has no stored diagonal, or the level search stalled (unsymmetric structure)" : "STS: bad arguments");
  // A' = P A P^T (reorderA): in HBM when the rows fit the warp sorter, else on the host
  int rc0 = ensure_stats(A);
  if (rc0) return rc0;
  bool done = false;
  if (A->stats[8] <= hspmv::BANDK_MAX_ROW) {
    int *rp2 = nullptr, *ci2 = nullptr;
    void *val2 = nullptr;
    u32 *d_perm = nullptr;
    if (cudaMalloc(&rp2, (size_t)(A->m + 1 + PAD) * 4) == cudaSuccess && cudaMalloc(&ci2, (size_t)(A->nnz + PAD) * 4) == cudaSuccess &&
        cudaMalloc(&val2, (size_t)(A->nnz + PAD) * vs) == cudaSuccess && cudaMalloc(&d_perm, (size_t)(A->m + PAD) * 4) == cudaSuccess) {
      cudaMemsetAsync((char *)rp2 + (size_t)(A->m + 1) * 4, 0, PAD * 4, A->stream);
      cudaMemsetAsync((char *)ci2 + (size_t)A->nnz * 4, 0, PAD * 4, A->stream);
      cudaMemsetAsync((char *)val2 + (size_t)A->nnz * vs, 0, PAD * vs, A->stream);
      cudaMemcpyAsync(d_perm, perm.data(), (size_t)A->m * 4, cudaMemcpyHostToDevice, A->stream);
      int64_t launches = 0;
      rc = hspmv::reorder_device(A->dtype, A->m, A->nnz, d_perm, (const u32 *)A->d_rp, (const u32 *)A->d_ci, A->d_val, (u32 *)rp2, (u32 *)ci2, val2,
                                 A->dev.sms, A->stream, &launches);
      g_launches.fetch_add(launches, std::memory_order_relaxed);
      if (!rc) {
        cudaFree(A->d_rp);
        cudaFree(A->d_ci);
        cudaFree(A->d_val);
        A->d_rp = rp2;
        A->d_ci = ci2;
        A->d_val = val2;
        A->d_perm = d_perm;
        CU(cudaMemcpy(A->h_rp.data(), A->d_rp, (size_t)(A->m + 1) * 4, cudaMemcpyDeviceToHost));
        ivec().swap(A->h_ci);
        std::vector<unsigned char>().swap(A->h_val);
        done = true;
      }
    }
    if (!done) {
      cudaFree(rp2);
      cudaFree(ci2);
      cudaFree(val2);
      cudaFree(d_perm);
      cudaGetLastError();
    }
  }
  if (!done) {
    if (A->h_val.size() != (size_t)A->nnz * vs) {
      A->h_val.resize((size_t)A->nnz * vs);
      if (A->nnz) CU(cudaMemcpy(A->h_val.data(), A->d_val, (size_t)A->nnz * vs, cudaMemcpyDeviceToHost));
    }
    ivec rp2((size_t)A->m + 1), ci2((size_t)A->nnz);
    std::vector<unsigned char> val2((size_t)A->nnz * vs);
    if (A->dtype == HSPMV_F64)
      hspmv::reorder_f64(A->m, perm.data(), (const u32 *)A->h_rp.data(), (const u32 *)A->h_ci.data(), (const double *)A->h_val.data(),
                         (u32 *)rp2.data(), (u32 *)ci2.data(), (double *)val2.data());
    else
      hspmv::reorder_f32(A->m, perm.data(), (const u32 *)A->h_rp.data(), (const u32 *)A->h_ci.data(), (const float *)A->h_val.data(),
                         (u32 *)rp2.data(), (u32 *)ci2.data(), (float *)val2.data());
    A->h_rp.swap(rp2);
    ivec().swap(A->h_ci);
    std::vector<unsigned char>().swap(A->h_val);
    CU(cudaMemcpy(A->d_rp, A->h_rp.data(), (size_t)(A->m + 1) * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(A->d_ci, ci2.data(), (size_t)A->nnz * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(A->d_val, val2.data(), (size_t)A->nnz * vs, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&A->d_perm, (size_t)(A->m + PAD) * 4));
    CU(cudaMemcpy(A->d_perm, perm.data(), (size_t)A->m * 4, cudaMemcpyHostToDevice));
  }
  A->h_perm.swap(perm);
  A->h_pack.swap(pack);
  A->n_packs = (int64_t)A->h_pack.size() - 1;
  // leading packs that one CTA of 1024 threads finishes in ONE pass run on CTA 0 alone (a block barrier per pack, ~2 us);
  // larger packs are worth a grid barrier (~3.5 us).  (r2 measurements, profiles/r2_sts_vs_reference_*.jsonl)
  const int64_t solo_rows = 1024 / sts_vec(A);
  A->graph_iters = 0; // a captured graph points at the old arrays
  A->solo_packs = 0;
  while (A->solo_packs < A->n_packs && A->h_pack[(size_t)A->solo_packs + 1] - A->h_pack[(size_t)A->solo_packs] <= solo_rows) ++A->solo_packs;
  A->bins_ready = false;
  A->merge_tile = 0;
  A->block_cap = 0;
  free_pipe(A);
  hspmv::row_stats(A->m, A->h_rp.data(), A->stats);
  A->stats_ready = true;
  CU(cudaMalloc(&A->d_diag, (size_t)(A->m + PAD) * 4));
  CU(cudaMalloc(&A->d_pack, (size_t)(A->n_packs + 1 + PAD) * 4));
  CU(cudaMalloc(&A->d_ready, (size_t)(A->m + 1 + PAD) * 4));
  CU(cudaMemset(A->d_ready, 0, (size_t)(A->m + 1 + PAD) * 4));
  CU(cudaMalloc(&A->d_sts_b, (size_t)(A->m + PAD) * vs));
  CU(cudaMalloc(&A->d_sts_x, (size_t)(A->m + PAD) * vs));
  CU(cudaMemset(A->d_sts_b, 0, (size_t)(A->m + PAD) * vs));
  CU(cudaMemset(A->d_sts_x, 0, (size_t)(A->m + PAD) * vs));
  CU(cudaMemcpy(A->d_pack, A->h_pack.data(), (size_t)(A->n_packs + 1) * 4, cudaMemcpyHostToDevice));
  int *d_bad = nullptr;
  CU(cudaMalloc(&d_bad, 4));
  CU(cudaMemset(d_bad, 0, 4));
  hspmv::sts_diag_kernel<<<grid_for(A->m, 256, A->dev.sms * 16), 256, 0, A->stream>>>(A->m, A->d_rp, A->d_ci, A->d_diag, d_bad);
  LAUNCHED();
  int bad = 0;
  CU(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, A->stream));
  CU(cudaStreamSynchronize(A->stream));
  cudaFree(d_bad);
  if (bad) return fail(HSPMV_EUNSUPPORTED, "STS: a row of the reordered matrix has no diagonal entry");
  A->sts_ready = true;
  return HSPMV_OK;
}

int hspmv_sts_get(const hspmv_matrix *A, uint32_t *perm, int32_t *pack_ptr, int64_t *n_packs) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  if (!A->sts_ready) return fail(HSPMV_ESTATE, "hspmv_build_sts first");
  if (perm) std::memcpy(perm, A->h_perm.data(), A->h_perm.size() * 4);
  if (pack_ptr) std::memcpy(pack_ptr, A->h_pack.data(), A->h_pack.size() * 4);
  if (n_packs) *n_packs = A->n_packs;
  return HSPMV_OK;
}

int hspmv_sts_compute_b(hspmv_matrix *A) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  if (!A->sts_ready) return fail(HSPMV_ESTATE, "hspmv_build_sts first");
  const int grid = grid_for(A->m, 256, A->dev.sms * 16);
  if (A->dtype == HSPMV_F64)
    hspmv::sts_b_kernel<double><<<grid, 256, 0, A->stream>>>(A->m, A->d_rp, A->d_diag, (const double *)A->d_val, (double *)A->d_sts_b);
  else
    hspmv::sts_b_kernel<float><<<grid, 256, 0, A->stream>>>(A->m, A->d_rp, A->d_diag, (const float *)A->d_val, (float *)A->d_sts_b);
  LAUNCHED();
  CU(cudaStreamSynchronize(A->stream));
  return HSPMV_OK;
}

// b / x cross the boundary in the solver's (reordered) numbering (original_order == 0, what the reference's arrays
// hold) or in the matrix's original numbering (the permutation is applied inside)
static int sts_vector_io(hspmv_matrix *A, void *dev, void *host, int original_order, bool upload) {
  const size_t vs = A->vsize();
  const int grid = grid_for(A->m, 256, A->dev.sms * 16);
  if (upload) {
    if (original_order) { // dev[i] = host[perm[i]]
      CU(cudaMemcpyAsync(A->d_tmp, host, (size_t)A->m * vs, cudaMemcpyHostToDevice, A->stream));
      if (A->dtype == HSPMV_F64) hspmv::gather_kernel<double><<<grid, 256, 0, A->stream>>>(A->m, A->d_perm, (const double *)A->d_tmp, (double *)dev);
      else hspmv::gather_kernel<float><<<grid, 256, 0, A->stream>>>(A->m, A->d_perm, (const float *)A->d_tmp, (float *)dev);
      LAUNCHED();
    } else {
      CU(cudaMemcpyAsync(dev, host, (size_t)A->m * vs, cudaMemcpyHostToDevice, A->stream));
    }
  } else {
    if (A->last_stream_valid && A->last_stream != A->stream) CU(cudaStreamSynchronize(A->last_stream));
    if (original_order) { // host[perm[i]] = dev[i]
      if (A->dtype == HSPMV_F64) hspmv::scatter_kernel<double><<<grid, 256, 0, A->stream>>>(A->m, A->d_perm, (const double *)dev, (double *)A->d_tmp);
      else hspmv::scatter_kernel<float><<<grid, 256, 0, A->stream>>>(A->m, A->d_perm, (const float *)dev, (float *)A->d_tmp);
      LAUNCHED();
      CU(cudaMemcpyAsync(host, A->d_tmp, (size_t)A->m * vs, cudaMemcpyDeviceToHost, A->stream));
    } else {
      CU(cudaMemcpyAsync(host, dev, (size_t)A->m * vs, cudaMemcpyDeviceToHost, A->stream));
    }
  }
  CU(cudaStreamSynchronize(A->stream));
  return HSPMV_OK;
}
int hspmv_sts_set_b(hspmv_matrix *A, const void *b_host, int original_order) {
  if (!A || !b_host) return fail(HSPMV_EINVAL, "null");
  if (!A->sts_ready) return fail(HSPMV_ESTATE, "hspmv_build_sts first");
  return sts_vector_io(A, A->d_sts_b, const_cast<void *>(b_host), original_order, true);
}
int hspmv_sts_get_b(hspmv_matrix *A, void *b_host, int original_order) {
  if (!A || !b_host) return fail(HSPMV_EINVAL, "null");
  if (!A->sts_ready) return fail(HSPMV_ESTATE, "hspmv_build_sts first");
  return sts_vector_io(A, A->d_sts_b, b_host, original_order, false);
}
int hspmv_sts_get_x(hspmv_matrix *A, void *x_host, int original_order) {
  if (!A || !x_host) return fail(HSPMV_EINVAL, "null");
  if (!A->sts_ready) return fail(HSPMV_ESTATE, "hspmv_build_sts first");
  return sts_vector_io(A, A->d_sts_x, x_host, original_order, false);
}

int hspmv_sts_solve(hspmv_matrix *A, void *stream) {
  if (!A) return fail(HSPMV_EINVAL, "null handle");
  if (!A->sts_ready) return fail(HSPMV_ESTATE, "hspmv_build_sts first");
  cudaStream_t st = stream ? (cudaStream_t)stream : A->stream;
  A->last_stream = st;
  A->last_stream_valid = true;
  // Two forms.  Level-scheduled: a block / grid barrier per pack -- good when packs are few and large.  Barrier-free:
  // per-row ready flags -- good when packs are many and small (12 290 packs of ~120 rows: 36.7 -> 4.4 ms), worse when a
  // pack is thousands of rows (cfg2: 5.8 vs 2.6 ms; most resident groups then spin on the previous pack).
  // HSPMV_OPT_STS_BARRIERS: 0 = choose by the mean pack size, 1 = level-scheduled, 2 = barrier-free.
  const int64_t mode = A->opt[HSPMV_OPT_STS_BARRIERS];
  const bool flags = mode == 2 || (mode == 0 && A->n_packs > 0 && A->m / A->n_packs < 256);
  if (!flags) return A->dtype == HSPMV_F64 ? launch_sts_solve<double>(A, st) : launch_sts_solve<float>(A, st);
  return A->dtype == HSPMV_F64 ? launch_sts_flags<double>(A, st) : launch_sts_flags<float>(A, st);
}

int hspmv_sts_status(hspmv_matrix *A, int *timed_out) {
  if (!A || !timed_out) return fail(HSPMV_EINVAL, "null");
  if (!A->sts_ready) return fail(HSPMV_ESTATE, "hspmv_build_sts first");
  if (A->last_stream_valid) CU(cudaStreamSynchronize(A->last_stream));
  CU(cudaMemcpy(timed_out, A->d_ready + A->m, 4, cudaMemcpyDeviceToHost));
  return HSPMV_OK;
}

int hspmv_sts_solve_timed(hspmv_matrix *A, int warmup, int iters, float *ms_out) {
  if (!A || iters < 1 || !ms_out) return fail(HSPMV_EINVAL, "bad arguments");
  for (int i = 0; i < warmup; ++i) {
    int rc = hspmv_sts_solve(A, A->stream);
    if (rc) return rc;
  }
  std::vector<cudaEvent_t> ev((size_t)iters * 2);
  for (auto &e : ev) CU(cudaEventCreate(&e));
  for (int i = 0; i < iters; ++i) {
    CU(cudaEventRecord(ev[(size_t)2 * i], A->stream));
    int rc = hspmv_sts_solve(A, A->stream);
    if (rc) return rc;
    CU(cudaEventRecord(ev[(size_t)2 * i + 1], A->stream));
  }
  CU(cudaStreamSynchronize(A->stream));
  for (int i = 0; i < iters; ++i) CU(cudaEventElapsedTime(&ms_out[i], ev[(size_t)2 * i], ev[(size_t)2 * i + 1]));
  for (auto &e : ev) cudaEventDestroy(e);
  return HSPMV_OK;
}

} // extern "C"
