// CSR-k size / launch-shape heuristics.  arch 0..2 reproduce the reference's fitted formulas so
// that the drop-in `spmv-auto-*` drivers print the same "using ssrs %d, srs %d" line:
//   0 Volta   cuda-spmv-csrk/cuda/spmv-auto-volta.cu:130-159 (= reformat-csr-to-csr3/spmv-auto.cpp:154-173,
//             helpers/csr3params.m:8-37)
//   1 Ampere  cuda-spmv-csrk/cuda/spmv-auto-ampere.cu:130-158
//   2 MI100   cuda-spmv-csrk/hip/spmv-auto-mi100.cu:130-159 (own base formula, Ampere-style rescale)
//   3 B200    this library: a super-row fills one shared-memory stage of the TMA kernel (about
//             432 nonzeros, a multiple of 32 / lanes-per-row rows); a super-super-row is one
//             CTA's worth of super-rows (16 warps).  Retuned from the r1 sweeps in profiles/.
#include "host_index.h"

#include <cmath>

namespace hspmv {

int auto_params(int64_t nnz, int64_t n_rows, int arch, int sizes[2], int *veclevel, int *block_dim_y) {
  if (n_rows <= 0 || nnz < 0 || !sizes) return HSPMV_EINVAL;
  const double d = (double)nnz / (double)n_rows;
  const double ld = std::log(d > 0 ? d : 1.0);
  int ssrs = 0, srs = 0, vec = 0, bdy = 12;
  if (arch == 0) {
    ssrs = (int)std::floor(8.89888 - 1.25 * ld + 0.5);
    srs = (int)std::floor(10.14618 - 1.5 * ld + 0.5);
    if (d > 8.0 && d <= 16.0) {
      vec = 4;
      ssrs = (int)std::floor((double)ssrs * 1.5 + 0.5);
      srs = ssrs * 2;
    } else if (d > 16.0 && d <= 32.0) {
      vec = 8; bdy = 8;
      ssrs *= 4;
      srs = ssrs >> 1;
    } else if (d > 32.0 && d <= 64.0) {
      vec = 16; bdy = 4;
      ssrs *= 5;
      srs = ssrs >> 1;
    } else if (d > 64.0) {
      vec = 32; bdy = 2;
      ssrs *= 5;
      srs = ssrs >> 1;
    }
  } else if (arch == 1 || arch == 2) {
    if (arch == 1) {
      ssrs = (int)std::floor(0.5 + (9.175 - 1.32 * ld));
      srs = (int)std::floor(0.5 + (20.5 - 3.5 * ld));
    } else {
      ssrs = (int)std::floor(0.5 + (8.489 - 1.15 * ld));
      srs = (int)std::floor(0.5 + (10.711 - 1.607 * ld));
    }
    if (d > 8.0 && d <= 16.0) {
      vec = 4;
      srs = ssrs * 4;
    } else if (d > 16.0 && d <= 32.0) {
      vec = 8; bdy = 8;
      ssrs = (int)std::floor((double)ssrs * 2.5 + 0.5);
      srs = ssrs * 3;
    } else if (d > 32.0 && d <= 64.0) {
      vec = 16; bdy = 4;
      ssrs *= 2;
      srs = ssrs * 2;
    } else if (d > 64.0) {
      vec = 32; bdy = 2;
      ssrs = (int)std::floor((double)ssrs * 2.7 + 0.5);
      srs = (int)std::floor((double)ssrs / 4 + 0.5);
    }
  } else if (arch == 3) {
    // Same rule as the kernel's own defaults (abi.cu:run_csrk): lanes per row = smallest power of
    // two that leaves each lane <= 14 nonzeros of a typical row (one unmasked chunk of gathers);
    // a super-row = whole passes of the warp holding ~432 nonzeros (16 warps x 2 shared-memory
    // stages of that fill an SM); a super-super-row = one CTA's 16 warps.
    vec = 1;
    while (vec < 32 && d / vec > 14.0) vec <<= 1;
    const int rows_per_pass = 32 / vec;
    int passes = (int)std::floor(432.0 / ((d > 1 ? d : 1.0) * rows_per_pass) + 0.5);
    if (passes < 1) passes = 1;
    ssrs = passes * rows_per_pass; // sizes[0]: rows per super-row (nnz-weighted) -- the reference calls it "ssrs" (Q9)
    srs = 16;                      // sizes[1]: super-rows per super-super-row = warps per CTA
    bdy = 16;
  } else {
    return HSPMV_EINVAL;
  }
  if (ssrs < 1) ssrs = 1;
  if (srs < 1) srs = 1;
  sizes[0] = ssrs;
  sizes[1] = srs;
  if (veclevel) *veclevel = vec;
  if (block_dim_y) *block_dim_y = bdy;
  return HSPMV_OK;
}

} // namespace hspmv
