// Instantiations of csrk_tma_kernel compiled for 640 threads per CTA (18 warps at 96 registers), with and without the
// halo support, for the shapes listed in HSPMV_CSRK_WIDE_VARIANTS (see csrk_variants.h).
#include "csrk_variants.h"

namespace hspmv {

template <typename T>
int launch_csrk_wide(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device, const CsrkVariant &v) {
  const bool halo = p.n_ranges > 1 || p.n_wait > 0;
#define X(vv, mm, tt)                                                                                        \
  if (v.vec == vv && v.main == mm && v.tail == tt)                                                           \
    return halo ? launch_csrk_one<T, vv, mm, tt, true, true>(p, grid, warps, smem, st, device)               \
                : launch_csrk_one<T, vv, mm, tt, false, true>(p, grid, warps, smem, st, device);
  HSPMV_CSRK_WIDE_VARIANTS(X)
#undef X
  return -1;
}

template int launch_csrk_wide<float>(const CsrkParams<float> &, int, int, size_t, cudaStream_t, int, const CsrkVariant &);
template int launch_csrk_wide<double>(const CsrkParams<double> &, int, int, size_t, cudaStream_t, int, const CsrkVariant &);

} // namespace hspmv
