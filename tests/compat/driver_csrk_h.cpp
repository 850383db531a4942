// A driver written against the reference's CSRk_Graph (spmv-csrk/csrk.h:309-338), following
// spmv-csrk/spmv.cpp:139-211: constructor from malloc'ed CSR arrays, putInCSRkFormat, setX,
// SpMV(y), getPermutation, check y[i] against the serial product at permutation[i].
// Compiled against include/csrk.h with -DHSPMV_VAL_DOUBLE (fp64) or without (the reference's fp32).
#include <cmath>
#include <cstring>
#include "csrk.h"

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  int64_t m = 0, n = 0, nnz = 0;
  int32_t *rp = NULL, *ci = NULL;
  void *val = NULL;
  if (hspmv_read_csr_text(argv[1], 1, HSPMV_VAL_DTYPE, &m, &n, &nnz, &rp, &ci, &val)) return 3;   // 1-based, spmv.cpp:60,67
  int k = std::atoi(argv[2]);
  int supRowSizes[2] = {std::atoi(argv[3]), argc > 4 ? std::atoi(argv[4]) : 8};
  val_type *v = (val_type *)val;
  std::vector<val_type> x((size_t)m), y((size_t)m, 0), yser((size_t)m, 0);
  for (int64_t i = 0; i < m; i++) x[(size_t)i] = (val_type)(1.0 + (i % 7) / 8.0);
  for (int64_t r = 0; r < m; r++)
    for (int j = rp[r]; j < rp[r + 1]; j++) yser[(size_t)r] += v[j] * x[(size_t)ci[j]];

  CSRk_Graph A_mat(m, m, nnz, (unsigned int *)rp, (unsigned int *)ci, v, "SpMV", "", "HAND", false, k, supRowSizes);
  A_mat.putInCSRkFormat();
  A_mat.setX(x.data());
  val_type *yp = y.data();
  A_mat.SpMV(yp);
  unsigned int *permutation = A_mat.getPermutation();
  int wrong = 0;
  for (int64_t i = 0; i < m; i++)
    if (std::fabs((double)y[(size_t)i] - (double)yser[permutation[i]]) > 0.01) wrong++;     // the reference's own tolerance (Q10)
  // CUDA-flavour calls: setY / SpMV() / getY
  std::vector<val_type> y2((size_t)m, (val_type)-1);
  A_mat.setY(y2.data());
  A_mat.SpMV();
  val_type *got = A_mat.getY();
  for (int64_t i = 0; i < m; i++)
    if (got[i] != y[(size_t)i]) wrong++;
  std::printf("size %d %d %d\ncoarsest rows %ld\nNumber Wrong: %d \n", (int)m, (int)n, (int)nnz, A_mat.getNumCoarsestRows(), wrong);
  hspmv_free_host(rp); hspmv_free_host(ci); hspmv_free_host(val);
  return 0;
}
