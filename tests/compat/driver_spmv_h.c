/* A driver written against the reference's spmv-csr/spmv.h:8-12 (my_read_csr + omp_spmv with the
 * double signatures), following the call sequence of spmv-csr/spmv.c:116-200: read, x = 1,
 * serial check product, timed product, count mismatches.  Compiled against include/spmv.h. */
#include <math.h>
#include "spmv.h"

int main(int argc, char **argv) {
  int m = 0, n = 0, nnz = 0, *row_start = NULL, *col_idx = NULL;
  double *val = NULL;
  if (argc < 2) return 2;
  my_read_csr(argv[1], &m, &n, &nnz, &row_start, &col_idx, &val);
  if (!m) return 3;
  double *x = (double *)malloc(sizeof(double) * (size_t)m), *y = (double *)calloc((size_t)m, sizeof(double));
  for (int i = 0; i < m; i++) x[i] = 1.0 + (i % 97) / 97.0;
  for (int it = 0; it < 3; it++) omp_spmv(m, n, nnz, row_start, col_idx, val, x, y);
  int wrong = 0;
  for (int r = 0; r < m; r++) {
    double s = 0, a = 0;
    for (int j = row_start[r]; j < row_start[r + 1]; j++) {
      s += val[j] * x[col_idx[j]];
      a += fabs(val[j] * x[col_idx[j]]);
    }
    if (fabs(s - y[r]) > 1e-12 * a) wrong++;
  }
  printf("size %d %d %d\nNumber Wrong: %d \n", m, n, nnz, wrong);
  hspmv_omp_spmv_cache_clear();
  free(row_start); free(col_idx); free(val); free(x); free(y);
  return 0;
}
