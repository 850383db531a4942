// A driver written the way cuda-spmv-csrk/cuda/spmv.cu:111-233 and spmv-auto-ampere.cu:200-220 are: it declares the
// library's device-pointer globals `extern`, builds a CSRk_Graph, and launches cuSpMV_2 / cuSpMV_3 / cuSpMV_3_vec BY
// NAME with the reference's launch geometries, then checks y against the serial product.  Compiled against
// include/csrk.cuh (tests/test_compat_headers.py).
//   usage: driver_cuspmv <1-based .csr file> <k> <size0> <size1>
#include <cmath>
#include <cstring>
#include "csrk.cuh"

extern unsigned int *N_gpu;
extern unsigned int *M_gpu;
extern unsigned int *NNZ_gpu;
extern unsigned int *numCoarsestRows_gpu;
extern unsigned int *mapCoarseToFinerRows_gpu_outer;
extern unsigned int *mapCoarseToFinerRows_gpu_inner;
extern unsigned int *r_vec_gpu;
extern unsigned int *c_vec_gpu;
extern val_type *val_gpu;
extern val_type *x_test_gpu;
extern val_type *y_gpu;
extern unsigned int *coarseRowIndex;

static int check(CSRk_Graph &A_mat, const std::vector<val_type> &yser, std::vector<val_type> &y, const char *what) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    std::printf("%s: %s\n", what, cudaGetErrorString(e));
    return 1 << 20;
  }
  val_type *got = A_mat.getY();
  unsigned int *permutation = A_mat.getPermutation();
  int wrong = 0;
  for (size_t i = 0; i < y.size(); i++) {
    const double want = (double)yser[permutation ? permutation[i] : i];
    if (!(std::fabs((double)got[i] - want) <= 1e-5 * (1.0 + std::fabs(want)))) wrong++;
  }
  std::printf("%s wrong: %d\n", what, wrong);
  // poison y for the next launch: a kernel that writes nothing must not pass
  std::vector<val_type> nanv(y.size(), (val_type)NAN);
  A_mat.setY(nanv.data());
  A_mat.setY(y.data());
  cudaMemset(y_gpu, 0xff, y.size() * sizeof(val_type));
  return wrong;
}

int main(int argc, char **argv) {
  if (argc < 5) return 2;
  int64_t m = 0, n = 0, nnz = 0;
  int32_t *rp = NULL, *ci = NULL;
  void *val = NULL;
  if (hspmv_read_csr_text(argv[1], 1, HSPMV_VAL_DTYPE, &m, &n, &nnz, &rp, &ci, &val)) return 3;
  int k = std::atoi(argv[2]);
  int supRowSizes[2] = {std::atoi(argv[3]), std::atoi(argv[4])};
  val_type *v = (val_type *)val;
  std::vector<val_type> x((size_t)m), y((size_t)m, 0), yser((size_t)m, 0);
  for (int64_t i = 0; i < m; i++) x[(size_t)i] = (val_type)(1.0 + (i % 7) / 8.0);
  for (int64_t r = 0; r < m; r++)
    for (int j = rp[r]; j < rp[r + 1]; j++) yser[(size_t)r] += v[j] * x[(size_t)ci[j]];

  CSRk_Graph A_mat(m, m, nnz, (unsigned int *)rp, (unsigned int *)ci, v, "SpMV", "", "HAND", false, k, supRowSizes);
  A_mat.putInCSRkFormat();
  A_mat.setX(x.data());
  A_mat.setY(y.data());
  cudaMemset(y_gpu, 0xff, (size_t)m * sizeof(val_type));
  int wrong = 0;
  const int blockDimx = 8, blockDimy = 12;
  if (k == 2) {
    cuSpMV_2<<<A_mat.getNumCoarsestRows(), dim3(blockDimx, blockDimy)>>>(numCoarsestRows_gpu, mapCoarseToFinerRows_gpu_outer, r_vec_gpu,
                                                                         c_vec_gpu, val_gpu, x_test_gpu, y_gpu);
    wrong += check(A_mat, yser, y, "cuSpMV_2 <<<nSR, (8,12)>>>");
    cuSpMV_2<<<7, dim3(5, 7)>>>(numCoarsestRows_gpu, mapCoarseToFinerRows_gpu_outer, r_vec_gpu, c_vec_gpu, val_gpu, x_test_gpu, y_gpu);
    wrong += check(A_mat, yser, y, "cuSpMV_2 <<<7, (5,7)>>> (strided grid, partial warp)");
    cuSpMV_2<<<A_mat.getNumCoarsestRows(), 256>>>(numCoarsestRows_gpu, mapCoarseToFinerRows_gpu_outer, r_vec_gpu, c_vec_gpu, val_gpu,
                                                  x_test_gpu, y_gpu);
    wrong += check(A_mat, yser, y, "cuSpMV_2 <<<nSR, 256>>>");
  } else {
    cuSpMV_3<<<A_mat.getNumCoarsestRows(), dim3(blockDimx, blockDimy)>>>(numCoarsestRows_gpu, mapCoarseToFinerRows_gpu_outer,
                                                                         mapCoarseToFinerRows_gpu_inner, r_vec_gpu, c_vec_gpu, val_gpu,
                                                                         x_test_gpu, y_gpu);
    wrong += check(A_mat, yser, y, "cuSpMV_3 <<<nSSR, (8,12)>>>");
    for (int veclevel = 4; veclevel <= 32; veclevel *= 2) {   // spmv-auto-ampere.cu:141-158: (4,8,12) (8,8,8) (16,8,4) (32,8,2)
      const int by = veclevel == 4 ? 12 : veclevel == 8 ? 8 : veclevel == 16 ? 4 : 2;
      cuSpMV_3_vec<<<A_mat.getNumCoarsestRows(), dim3(veclevel, blockDimx, by)>>>(numCoarsestRows_gpu, mapCoarseToFinerRows_gpu_outer,
                                                                                  mapCoarseToFinerRows_gpu_inner, r_vec_gpu, c_vec_gpu,
                                                                                  val_gpu, x_test_gpu, y_gpu);
      wrong += check(A_mat, yser, y, "cuSpMV_3_vec <<<nSSR, (veclevel,8,by)>>>");
    }
  }
  unsigned int dims[4] = {0, 0, 0, 0};
  cudaMemcpy(dims, N_gpu, sizeof(dims), cudaMemcpyDeviceToHost);
  if (dims[0] != (unsigned)m || dims[1] != (unsigned)m || dims[2] != (unsigned)nnz || dims[3] != (unsigned)A_mat.getNumCoarsestRows() ||
      M_gpu != N_gpu + 1 || NNZ_gpu != N_gpu + 2 || numCoarsestRows_gpu != N_gpu + 3)
    wrong += 1 << 21;
  std::printf("size %d %d %d\ncoarsest rows %ld\nNumber Wrong: %d \n", (int)m, (int)n, (int)nnz, A_mat.getNumCoarsestRows(), wrong);
  hspmv_free_host(rp); hspmv_free_host(ci); hspmv_free_host(val);
  return 0;
}
