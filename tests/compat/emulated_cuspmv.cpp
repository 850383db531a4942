// The kernels of include/csrk.cuh (cuSpMV_2 / cuSpMV_3 / cuSpMV_3_vec) executed WITHOUT a GPU on the host emulation of
// emulate_cuda.h, with the reference's launch geometries (cuda-spmv-csrk/cuda/spmv.cu:189-198,
// spmv-auto-ampere.cu:141-158,208-218), on a Band-k-reordered matrix built with the library's host builder.
//   usage: emulated_cuspmv <1-based .csr file> <k> <size0> <size1>
// Prints "Number Wrong: N" (rows of y that differ from the serial product beyond rounding, NaN-poisoned before every
// launch) and "Shuffle Errors: E" (a lane read a lane that had exited or skipped the shuffle).
#define HSPMV_COMPAT_HOST_EMULATION
#include "emulate_cuda.h"

#include <cmath>
#include "csrk.cuh"

int main(int argc, char **argv) {
  if (argc < 5) return 2;
  int64_t m = 0, n = 0, nnz = 0;
  int32_t *rp = NULL, *ci = NULL;
  void *val = NULL;
  if (hspmv_read_csr_text(argv[1], 1, HSPMV_VAL_DTYPE, &m, &n, &nnz, &rp, &ci, &val)) return 3;
  const int k = std::atoi(argv[2]);
  int sizes[2] = {std::atoi(argv[3]), std::atoi(argv[4])};
  std::vector<unsigned> perm((size_t)m), sr((size_t)m + 1), ssr((size_t)m + 1);
  int64_t n_sr = 0, n_ssr = 0;
  if (hspmv_bandk_host(m, nnz, (const uint32_t *)rp, (const uint32_t *)ci, k, sizes, perm.data(), sr.data(), &n_sr, ssr.data(), &n_ssr)) return 4;
  std::vector<unsigned> rp2((size_t)m + 1), ci2((size_t)nnz);
  std::vector<val_type> v2((size_t)nnz);
  if (hspmv_reorder_host(m, nnz, perm.data(), (const uint32_t *)rp, (const uint32_t *)ci, val, HSPMV_VAL_DTYPE, rp2.data(), ci2.data(), v2.data()))
    return 5;
  // x in the new numbering (setX), serial product of the REORDERED matrix as the expected y
  std::vector<val_type> x((size_t)m), y((size_t)m), want((size_t)m, 0);
  for (int64_t i = 0; i < m; i++) x[(size_t)i] = (val_type)(1.0 + (perm[(size_t)i] % 7) / 8.0);
  for (int64_t r = 0; r < m; r++)
    for (unsigned j = rp2[(size_t)r]; j < rp2[(size_t)r + 1]; j++) want[(size_t)r] += v2[j] * x[ci2[j]];
  unsigned coarsest = (unsigned)(k == 3 ? n_ssr : n_sr);
  unsigned *outer = k == 3 ? ssr.data() : sr.data(), *inner = sr.data();

  int wrong = 0, launches = 0;
  auto check = [&](const char *what) {
    int w = 0;
    for (int64_t i = 0; i < m; i++)
      if (!(std::fabs((double)y[(size_t)i] - (double)want[(size_t)i]) <= 1e-5 * (1.0 + std::fabs((double)want[(size_t)i])))) w++;
    std::printf("%s wrong: %d\n", what, w);
    wrong += w;
    ++launches;
    std::fill(y.begin(), y.end(), (val_type)NAN);
  };
  std::fill(y.begin(), y.end(), (val_type)NAN);
  if (k == 2) {
    emu::launch(cuSpMV_2, dim3(coarsest), dim3(8, 12), &coarsest, outer, rp2.data(), ci2.data(), v2.data(), x.data(), y.data());
    check("cuSpMV_2 <<<nSR, (8,12)>>>");
    emu::launch(cuSpMV_2, dim3(7), dim3(5, 7), &coarsest, outer, rp2.data(), ci2.data(), v2.data(), x.data(), y.data());
    check("cuSpMV_2 <<<7, (5,7)>>>");
    emu::launch(cuSpMV_2, dim3(coarsest / 3 + 1), dim3(64), &coarsest, outer, rp2.data(), ci2.data(), v2.data(), x.data(), y.data());
    check("cuSpMV_2 <<<nSR/3+1, 64>>>");
  } else {
    emu::launch(cuSpMV_3, dim3(coarsest), dim3(8, 12), &coarsest, outer, inner, rp2.data(), ci2.data(), v2.data(), x.data(), y.data());
    check("cuSpMV_3 <<<nSSR, (8,12)>>>");
    for (int veclevel = 4; veclevel <= 32; veclevel *= 2) {
      const int by = veclevel == 4 ? 12 : veclevel == 8 ? 8 : veclevel == 16 ? 4 : 2;
      emu::launch(cuSpMV_3_vec, dim3(coarsest), dim3(veclevel, 8, by), &coarsest, outer, inner, rp2.data(), ci2.data(), v2.data(), x.data(),
                  y.data());
      check("cuSpMV_3_vec <<<nSSR, (veclevel,8,by)>>>");
    }
  }
  std::printf("launches %d shuffles %ld\nNumber Wrong: %d \nShuffle Errors: %d \n", launches, emu::shuffles.load(), wrong, emu::errors.load());
  hspmv_free_host(rp); hspmv_free_host(ci); hspmv_free_host(val);
  return 0;
}
