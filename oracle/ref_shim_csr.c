/* TEST INFRASTRUCTURE ONLY -- timing helper linked next to the UNMODIFIED reference
 * /root/reference/spmv-csr/spmv.c (which provides omp_spmv :92-114, test_spmv :68-90,
 * my_read_csr :11-57).  `float` is rewritten to `double` by ref_prelude_c.h in the f64 build. */
void omp_spmv(int m, int n, int nnz, int *row_start, int *col_idx, float *val, float *x,
              float *y);

/* 5 warm-ups + N timed calls, min time in seconds: the loop at spmv-csr/spmv.c:164-181 */
double ref_csr_spmv_timed(int m, int n, int nnz, int *row_start, int *col_idx, float *val,
                          float *x, float *y, int warmup, int iters, double *avg_out) {
  int i;
  double best = 1e30, sum = 0;
  for (i = 0; i < warmup; ++i) omp_spmv(m, n, nnz, row_start, col_idx, val, x, y);
  for (i = 0; i < iters; ++i) {
    double t0 = omp_get_wtime();
    omp_spmv(m, n, nnz, row_start, col_idx, val, x, y);
    double t = omp_get_wtime() - t0;
    if (t < best) best = t;
    sum += t;
  }
  if (avg_out) *avg_out = iters > 0 ? sum / iters : 0;
  return best;
}

int ref_csr_sizeof_real(void) { return (int)sizeof(float); }
