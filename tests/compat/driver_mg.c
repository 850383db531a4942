/* A plain C99 program in the style of the reference drivers that uses more than one GPU through the multi-GPU
 * entry points of include/hspmv.h (hspmv_mg_*): build a 27-point stencil on a g^3 grid on the host, shard it over
 * the devices given on the command line, run a few steps with a DIFFERENT x each step (halo regions poisoned with
 * NaN first), and check every y against a serial product.  usage: driver_mg <g> <n_dev> <k>   */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "hspmv.h"

#define CK(call)                                                     \
  do {                                                               \
    if (call) {                                                      \
      printf("Error: %s | %s\n", hspmv_mg_last_error(), hspmv_last_error()); \
      return 2;                                                      \
    }                                                                \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 4) {
    printf("usage: %s g n_dev k\n", argv[0]);
    return 0;
  }
  const int g = atoi(argv[1]), n_dev = atoi(argv[2]), k = atoi(argv[3]);
  const int64_t n = (int64_t)g * g * g;
  int32_t *rp = (int32_t *)malloc((size_t)(n + 1) * 4);
  int32_t *ci = (int32_t *)malloc((size_t)n * 27 * 4);
  double *val = (double *)malloc((size_t)n * 27 * 8);
  int64_t e = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int x = (int)(i % g), y = (int)((i / g) % g), z = (int)(i / ((int64_t)g * g));
    rp[i] = (int32_t)e;
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          if (z + dz < 0 || z + dz >= g || y + dy < 0 || y + dy >= g || x + dx < 0 || x + dx >= g) continue;
          const int64_t c = i + ((int64_t)dz * g + dy) * g + dx;
          ci[e] = (int32_t)c;
          val[e] = c == i ? 32.0 + (double)((7 * i) % 64) / 64.0 : -(1.0 + (double)((31 * i + 17 * c) % 64) / 64.0);
          ++e;
        }
  }
  rp[n] = (int32_t)e;
  int devices[64], sizes[2] = {16, 8};
  for (int d = 0; d < n_dev; ++d) devices[d] = d;
  hspmv_mg *M = NULL;
  CK(hspmv_mg_create_csr(&M, devices, n_dev, n, n, e, rp, ci, val, HSPMV_F64, 0, k, k > 1 ? sizes : NULL));
  int repl = 0;
  int64_t row_lo[65], halo[64];
  CK(hspmv_mg_info(M, NULL, NULL, NULL, NULL, NULL, NULL, &repl));
  CK(hspmv_mg_partition_query(M, NULL, row_lo, NULL, NULL, NULL, halo));
  printf("devices %d replicated %d rows/dev0 %lld halo0 %lld\n", n_dev, repl, (long long)(row_lo[1] - row_lo[0]), (long long)halo[0]);
  CK(hspmv_mg_poison_halos(M));
  double *x = (double *)malloc((size_t)n * 8), *y = (double *)malloc((size_t)n * 8);
  int wrong = 0;
  for (int t = 1; t <= 5; ++t) {
    for (int64_t i = 0; i < n; ++i) x[i] = 1.0 + (double)((i * (t + 3) + 11 * t) % 97) / 128.0;
    CK(hspmv_mg_set_x(M, x));
    CK(hspmv_mg_spmv(M));
    CK(hspmv_mg_get_y(M, y));
    for (int64_t r = 0; r < n; ++r) {
      double acc = 0, scale = 0;
      for (int64_t j = rp[r]; j < rp[r + 1]; ++j) {
        acc += val[j] * x[ci[j]];
        scale += fabs(val[j] * x[ci[j]]);
      }
      if (!(fabs(acc - y[r]) <= 1e-12 * scale)) ++wrong;
    }
  }
  float ms = 0;
  CK(hspmv_mg_spmv_timed_batch(M, 2, 5, &ms));
  printf("ms per step %g\n", (double)ms);
  printf("Number Wrong: %d \n", wrong);
  hspmv_mg_destroy(M);
  free(rp); free(ci); free(val); free(x); free(y);
  return wrong ? 1 : 0;
}
