// csrk.cuh -- CUDA flavour of the CSRk_Graph mirror: for drivers written like the reference's
// cuda-spmv-csrk/cuda/spmv.cu and spmv-auto-{volta,ampere}.cu, which do not call A_mat.SpMV() but launch the
// library's kernels BY NAME on the library's device-pointer globals:
//
//     cuSpMV_2<<<A_mat.getNumCoarsestRows(), dim3(blockDimx, blockDimy)>>>(numCoarsestRows_gpu,
//         mapCoarseToFinerRows_gpu_outer, r_vec_gpu, c_vec_gpu, val_gpu, x_test_gpu, y_gpu);     (spmv.cu:195-198)
//     cuSpMV_3 / cuSpMV_3_vec<<<nSSR, dim3([veclevel,] blockDimx, blockDimy)>>>(...)            (spmv.cu:189-193,
//                                                                                     spmv-auto-ampere.cu:208-218)
//
// This header keeps that source compiling and correct against libhspmv:
//   * the globals the drivers declare `extern` (spmv.cu:33-44; defined in the reference's csrk.cu) are defined here
//     (weak, so any number of translation units may include the header) and are pointed at the handle's device arrays
//     by the constructor and by putInCSRkFormat() through hspmv_device_view_get();
//   * cuSpMV_2 / cuSpMV_3 / cuSpMV_3_vec exist with the reference's parameter lists (csrk.cuh:26-48) and accept the
//     reference's launch geometries: grid = number of coarsest rows (cuSpMV_2 also strides when the grid is smaller),
//     any 1-, 2- or 3-D block.
// Unlike the reference (quirk Q3: it uploads the matrix BEFORE reordering it and then runs the post-RCM maps on it)
// the arrays behind the globals are the REORDERED matrix, x_test_gpu is x in the new numbering and y_gpu comes back
// in the new numbering (A_mat.getY()), i.e. y[i] = (A x)[perm[i]] as in the CPU driver (spmv-csrk/spmv.cpp:203).
//
// These kernels are a compatibility surface, not the fast path: the caller's geometry (one small CTA per coarsest
// row) rules out the persistent TMA tile engine.  Within that geometry they are written for sm_100a -- a coarsest
// row is a contiguous range of rows; the CTA's threads form groups of VEC lanes per row (VEC from the range's mean
// row length, so the column / value reads of a group are contiguous), partial sums meet in a shuffle tree -- instead
// of the reference's one row per thread (csrk.cu:320-391) or its 544-float shared-memory reduction (:184-242).
// For speed, replace the launch line by A_mat.SpMV(): same result, the library's own launch.
#ifndef HSPMV_COMPAT_CSRK_CUH
#define HSPMV_COMPAT_CSRK_CUH
#ifndef HSPMV_COMPAT_HOST_EMULATION // tests/compat/emulate_cuda.h runs the kernels below on host threads instead
#include <cuda_runtime.h>
#endif

struct hspmv_matrix;
static inline void hspmv_compat_bind_device_globals(hspmv_matrix *A);
#define HSPMV_CSRK_BIND_DEVICE_GLOBALS(A) hspmv_compat_bind_device_globals(A)
#include "csrk.h"

// ---- the reference's device-pointer globals (spmv.cu:33-44) -----------------------------------------------------
#define HSPMV_WEAK __attribute__((weak))
HSPMV_WEAK unsigned int *N_gpu = NULL;
HSPMV_WEAK unsigned int *M_gpu = NULL;
HSPMV_WEAK unsigned int *NNZ_gpu = NULL;
HSPMV_WEAK unsigned int *numCoarsestRows_gpu = NULL;
HSPMV_WEAK unsigned int *mapCoarseToFinerRows_gpu_outer = NULL;
HSPMV_WEAK unsigned int *mapCoarseToFinerRows_gpu_inner = NULL;
HSPMV_WEAK unsigned int *r_vec_gpu = NULL;
HSPMV_WEAK unsigned int *c_vec_gpu = NULL;
HSPMV_WEAK val_type *val_gpu = NULL;
HSPMV_WEAK val_type *x_test_gpu = NULL;
HSPMV_WEAK val_type *y_gpu = NULL;
HSPMV_WEAK unsigned int *coarseRowIndex = NULL; // declared by the drivers, never used by them

static inline void hspmv_compat_bind_device_globals(hspmv_matrix *A) {
  hspmv_device_view v;
  if (hspmv_device_view_get(A, &v)) {
    std::cout << "csrk.cuh: " << hspmv_last_error() << std::endl;
    std::exit(1);
  }
  unsigned int *dims = (unsigned int *)v.dims;
  N_gpu = dims;
  M_gpu = dims + 1;
  NNZ_gpu = dims + 2;
  numCoarsestRows_gpu = dims + 3;
  r_vec_gpu = (unsigned int *)v.row_ptr;
  c_vec_gpu = (unsigned int *)v.col_idx;
  val_gpu = (val_type *)v.val;
  x_test_gpu = (val_type *)v.x;
  y_gpu = (val_type *)v.y;
  // k = 2: the drivers hand `mapCoarseToFinerRows_gpu_outer` to cuSpMV_2 as its only map (spmv.cu:197);
  // k = 3: outer = super-super-rows -> super-rows, inner = super-rows -> rows (csrk.cu:247-262)
  mapCoarseToFinerRows_gpu_outer = (unsigned int *)(v.k == 3 ? v.ssr_ptr : v.sr_ptr);
  mapCoarseToFinerRows_gpu_inner = (unsigned int *)v.sr_ptr;
}

// ---- kernels ------------------------------------------------------------------------------------------------------
namespace hspmv_compat {

// rows [r0, r1) by a CTA of `nthreads` threads (flattened id `tid`), VEC lanes per row.  Every thread runs the same
// number of trips of the row loop (the shuffles below need whole warps); nthreads is a multiple of VEC.
template <int VEC>
__device__ __forceinline__ void rows(unsigned r0, unsigned r1, const unsigned *__restrict__ rp, const unsigned *__restrict__ ci,
                                     const val_type *__restrict__ val, const val_type *__restrict__ x, val_type *__restrict__ y,
                                     unsigned tid, unsigned nthreads) {
  const unsigned groups = nthreads / VEC, g = tid / VEC, lane = tid % VEC;
  for (unsigned base = r0; base < r1; base += groups) {
    const unsigned row = base + g;
    val_type a0 = 0, a1 = 0;
    if (row < r1) {
      const unsigned e = __ldg(rp + row + 1);
      unsigned j = __ldg(rp + row) + lane;
      for (; j + VEC < e; j += 2 * VEC) { // two independent gathers in flight per lane
        a0 += __ldg(val + j) * __ldg(x + __ldg(ci + j));
        a1 += __ldg(val + j + VEC) * __ldg(x + __ldg(ci + j + VEC));
      }
      if (j < e) a0 += __ldg(val + j) * __ldg(x + __ldg(ci + j));
    }
    val_type acc = a0 + a1;
#pragma unroll
    for (int o = VEC / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && row < r1) y[row] = acc;
  }
}

// lanes per row for a range of rows: smallest power of two that leaves a lane <= 4 entries of a mean row.  Whole
// warps are needed for the shuffle tree, so a block whose size is not a multiple of 32 runs one row per thread.
__device__ __forceinline__ void range(unsigned r0, unsigned r1, const unsigned *__restrict__ rp, const unsigned *__restrict__ ci,
                                      const val_type *__restrict__ val, const val_type *__restrict__ x, val_type *__restrict__ y) {
  if (r1 <= r0) return;
  const unsigned tid = (threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x;
  const unsigned nthreads = blockDim.x * blockDim.y * blockDim.z;
  const unsigned mean = (__ldg(rp + r1) - __ldg(rp + r0)) / (r1 - r0);
  if ((nthreads & 31u) || mean <= 4)
    rows<1>(r0, r1, rp, ci, val, x, y, tid, nthreads);
  else if (mean <= 8)
    rows<2>(r0, r1, rp, ci, val, x, y, tid, nthreads);
  else if (mean <= 16)
    rows<4>(r0, r1, rp, ci, val, x, y, tid, nthreads);
  else if (mean <= 32)
    rows<8>(r0, r1, rp, ci, val, x, y, tid, nthreads);
  else if (mean <= 64)
    rows<16>(r0, r1, rp, ci, val, x, y, tid, nthreads);
  else
    rows<32>(r0, r1, rp, ci, val, x, y, tid, nthreads);
}

} // namespace hspmv_compat

// k = 2 (csrk.cu:320-391): one CTA per super-row, striding when the grid is smaller than their number
static __global__ void cuSpMV_2(unsigned int *__restrict__ numCoarsestRows_gpu, unsigned int *__restrict__ mapCoarseToFinerRows_gpu,
                                unsigned int *__restrict__ r_vec_gpu, unsigned int *__restrict__ c_vec_gpu,
                                val_type *__restrict__ val_gpu, val_type *__restrict__ x_test_gpu, val_type *__restrict__ y_gpu) {
  const unsigned n = *numCoarsestRows_gpu;
  for (unsigned s = blockIdx.x; s < n; s += gridDim.x)
    hspmv_compat::range(mapCoarseToFinerRows_gpu[s], mapCoarseToFinerRows_gpu[s + 1], r_vec_gpu, c_vec_gpu, val_gpu, x_test_gpu, y_gpu);
}

// k = 3 (csrk.cu:244-319 and :184-242): one CTA per super-super-row; its super-rows are consecutive, so the CTA
// walks the row range [inner[outer[b]], inner[outer[b+1]]).  The reference's two variants differ in how the block is
// shaped ((x, y) vs (veclevel, x, y)); both shapes land in the same code here.
static __global__ void cuSpMV_3(unsigned int *__restrict__ numCoarsestRows_gpu, unsigned int *__restrict__ mapCoarseToFinerRows_gpu_outer,
                                unsigned int *__restrict__ mapCoarseToFinerRows_gpu_inner, unsigned int *__restrict__ r_vec_gpu,
                                unsigned int *__restrict__ c_vec_gpu, val_type *__restrict__ val_gpu,
                                val_type *__restrict__ x_test_gpu, val_type *__restrict__ y_gpu) {
  const unsigned n = *numCoarsestRows_gpu;
  for (unsigned b = blockIdx.x; b < n; b += gridDim.x)
    hspmv_compat::range(mapCoarseToFinerRows_gpu_inner[mapCoarseToFinerRows_gpu_outer[b]],
                        mapCoarseToFinerRows_gpu_inner[mapCoarseToFinerRows_gpu_outer[b + 1]], r_vec_gpu, c_vec_gpu, val_gpu,
                        x_test_gpu, y_gpu);
}

static __global__ void cuSpMV_3_vec(unsigned int *__restrict__ numCoarsestRows_gpu, unsigned int *__restrict__ mapCoarseToFinerRows_gpu_outer,
                                    unsigned int *__restrict__ mapCoarseToFinerRows_gpu_inner, unsigned int *__restrict__ r_vec_gpu,
                                    unsigned int *__restrict__ c_vec_gpu, val_type *__restrict__ val_gpu,
                                    val_type *__restrict__ x_test_gpu, val_type *__restrict__ y_gpu) {
  const unsigned n = *numCoarsestRows_gpu;
  for (unsigned b = blockIdx.x; b < n; b += gridDim.x)
    hspmv_compat::range(mapCoarseToFinerRows_gpu_inner[mapCoarseToFinerRows_gpu_outer[b]],
                        mapCoarseToFinerRows_gpu_inner[mapCoarseToFinerRows_gpu_outer[b + 1]], r_vec_gpu, c_vec_gpu, val_gpu,
                        x_test_gpu, y_gpu);
}

#endif
