// Instantiations of csrk_tma_kernel WITH the halo support (hspmv_spmv_halo), group 8 (see csrk_variants.h).
#include "csrk_variants.h"

namespace hspmv {

template <typename T>
int launch_csrk_halo_group8(const CsrkParams<T> &p, int grid, int warps, size_t smem, cudaStream_t st, int device,
                            const CsrkVariant &v) {
#define X(vv, mm, tt) \
  if (v.vec == vv && v.main == mm && v.tail == tt) return launch_csrk_one<T, vv, mm, tt, true>(p, grid, warps, smem, st, device);
  HSPMV_CSRK_VARIANTS_V8(X)
#undef X
  return -1;
}

template int launch_csrk_halo_group8<float>(const CsrkParams<float> &, int, int, size_t, cudaStream_t, int, const CsrkVariant &);
template int launch_csrk_halo_group8<double>(const CsrkParams<double> &, int, int, size_t, cudaStream_t, int, const CsrkVariant &);

} // namespace hspmv
