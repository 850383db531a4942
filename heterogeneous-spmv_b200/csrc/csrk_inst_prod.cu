// Instantiations of csrk_prod_kernel (see csrk_variants.h).
#include "csrk_variants.h"

namespace hspmv {

template <typename T, int VEC, int U>
static int launch_prod_one(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device) {
  auto kern = csrk_prod_kernel<T, VEC, U>;
  static size_t configured[64] = {0};
  size_t &have = configured[device & 63];
  if (smem > have) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    have = smem;
  }
  kern<<<grid, warps * 32, smem, st>>>(p);
  return (int)cudaGetLastError();
}

template <typename T>
int launch_csrk_prod(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device, int vec, int u) {
#define X(vv, uu) \
  if (vec == vv && u == uu) return launch_prod_one<T, vv, uu>(p, grid, warps, smem, st, device);
  HSPMV_CSRK_PROD_VARIANTS(X)
#undef X
  return -1;
}

template int launch_csrk_prod<float>(const CsrkParams<float> &, int, int, size_t, cudaStream_t, int, int, int);
template int launch_csrk_prod<double>(const CsrkParams<double> &, int, int, size_t, cudaStream_t, int, int, int);

} // namespace hspmv
