// Compiled (lanes per row, unmasked chunk, masked chunk) shapes of csrk_tma_kernel and the
// dispatcher that launches one of them.  The instantiations live in csrk_inst_*.cu so that the
// build can compile them in parallel.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace hspmv {

struct CsrkVariant {
  int vec, main, tail;
};

// X(vec, main, tail).  main = 0: fully masked loop (ragged rows: one round of `tail` gathers per
// lane covers vec * tail entries whatever the row length; the first entry of a (vec, main) pair
// is the default, HSPMV_OPT_TAIL selects another).
#define HSPMV_CSRK_VARIANTS_V1(X) \
  X(1, 0, 4) X(1, 0, 8) X(1, 0, 12) X(1, 3, 2) X(1, 4, 2) X(1, 5, 2) X(1, 6, 2) X(1, 7, 4) X(1, 8, 4) X(1, 9, 4) X(1, 10, 4) X(1, 11, 4) X(1, 12, 4) X(1, 13, 4)
#define HSPMV_CSRK_VARIANTS_V2(X) \
  X(2, 0, 8) X(2, 0, 12) X(2, 3, 2) X(2, 4, 2) X(2, 5, 2) X(2, 6, 2) X(2, 7, 2) X(2, 8, 2) X(2, 9, 2) X(2, 10, 1) X(2, 11, 1) X(2, 12, 1) X(2, 13, 1)
#define HSPMV_CSRK_VARIANTS_V4(X) \
  X(4, 0, 8) X(4, 0, 12) X(4, 0, 14) X(4, 3, 1) X(4, 4, 2) X(4, 5, 2) X(4, 6, 1) X(4, 7, 2) X(4, 8, 2) X(4, 9, 2) X(4, 10, 1) X(4, 11, 1) X(4, 12, 1) X(4, 13, 1)
#define HSPMV_CSRK_VARIANTS_V8(X) \
  X(8, 0, 4) X(8, 0, 8) X(8, 0, 12) X(8, 3, 1) X(8, 4, 2) X(8, 5, 2) X(8, 6, 2) X(8, 7, 2) X(8, 8, 2) X(8, 9, 2) X(8, 10, 1) X(8, 11, 1) X(8, 12, 1) X(8, 13, 1) \
  X(16, 0, 4) X(16, 0, 8) X(16, 4, 2) X(16, 8, 2) X(16, 12, 1) X(32, 0, 4) X(32, 4, 2) X(32, 8, 2) X(32, 12, 1)
#define HSPMV_CSRK_VARIANTS(X) \
  HSPMV_CSRK_VARIANTS_V1(X) HSPMV_CSRK_VARIANTS_V2(X) HSPMV_CSRK_VARIANTS_V4(X) HSPMV_CSRK_VARIANTS_V8(X)

static const CsrkVariant kCsrkVariants[] = {
#define X(v, m, t) {v, m, t},
    HSPMV_CSRK_VARIANTS(X)
#undef X
};
static const int kNumCsrkVariants = (int)(sizeof(kCsrkVariants) / sizeof(kCsrkVariants[0]));

// one group of variants per translation unit; returns 0, a cudaError_t, or -1 (not in this group)
template <typename T> int launch_csrk_group1(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device, const CsrkVariant &v);
template <typename T> int launch_csrk_group2(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device, const CsrkVariant &v);
template <typename T> int launch_csrk_group4(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device, const CsrkVariant &v);
template <typename T> int launch_csrk_group8(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device, const CsrkVariant &v);

// shapes that also exist compiled for 640 threads per CTA (csrk_inst_wide.cu): X(vec, main, tail)
#define HSPMV_CSRK_WIDE_VARIANTS(X) X(2, 13, 1) X(2, 12, 1)
enum { CSRK_WIDE_WARPS = 18 };
inline bool csrk_has_wide(const CsrkVariant &v) {
#define X(vv, mm, tt) \
  if (v.vec == vv && v.main == mm && v.tail == tt) return true;
  HSPMV_CSRK_WIDE_VARIANTS(X)
#undef X
  return false;
}
template <typename T> int launch_csrk_wide(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device, const CsrkVariant &v);

// the same groups with the halo support compiled in (csrk_inst_halo_*.cu)
template <typename T> int launch_csrk_halo_group1(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device, const CsrkVariant &v);
template <typename T> int launch_csrk_halo_group2(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device, const CsrkVariant &v);
template <typename T> int launch_csrk_halo_group4(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device, const CsrkVariant &v);
template <typename T> int launch_csrk_halo_group8(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device, const CsrkVariant &v);

template <typename T>
inline int launch_csrk_variant(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st,
                               int device, const CsrkVariant &v) {
  if (warps > 16) return launch_csrk_wide<T>(p, grid, warps, smem, st, device, v); // only the caller's wide shapes get here
  if (p.n_ranges > 1 || p.n_wait > 0) { // hspmv_spmv_halo
    if (v.vec == 1) return launch_csrk_halo_group1<T>(p, grid, warps, smem, st, device, v);
    if (v.vec == 2) return launch_csrk_halo_group2<T>(p, grid, warps, smem, st, device, v);
    if (v.vec == 4) return launch_csrk_halo_group4<T>(p, grid, warps, smem, st, device, v);
    return launch_csrk_halo_group8<T>(p, grid, warps, smem, st, device, v);
  }
  if (v.vec == 1) return launch_csrk_group1<T>(p, grid, warps, smem, st, device, v);
  if (v.vec == 2) return launch_csrk_group2<T>(p, grid, warps, smem, st, device, v);
  if (v.vec == 4) return launch_csrk_group4<T>(p, grid, warps, smem, st, device, v);
  return launch_csrk_group8<T>(p, grid, warps, smem, st, device, v);
}

template <typename T, int VEC, int MAIN, int TAIL, bool HALO = false, bool WIDE = false>
inline int launch_csrk_one(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device) {
  auto kern = csrk_tma_kernel<T, VEC, MAIN, TAIL, HALO, WIDE>;
  static size_t configured[64] = {0}; // per device: opt-in shared memory already granted
  size_t &have = configured[device & 63];
  if (smem > have) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    have = smem;
  }
  kern<<<grid, warps * 32, smem, st>>>(p);
  return (int)cudaGetLastError();
}

} // namespace hspmv
