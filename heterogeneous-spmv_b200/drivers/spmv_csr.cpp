// Drop-in for cuda-spmv-csr/spmv.exe (cuda-spmv-csr/spmv.cu:185-305) and, by CLI, for
// spmv-csr/spmv.exe:   spmv.exe <file.mtx.rcm.csr> <num_runs> [--fp32] [--base 0|1] [--algo ...]
// 0-based reader, x = 1, N timed launches, TimeMin/TimeMax/TimeAvg, serial self-check.
#include "common.h"

template <typename T> static int run(const Options &o, HostCsr &H, int N) {
  if (o.gpus > 1) return run_multi_gpu<T>(o, H, N, /*warmup=*/0, 1, nullptr, 0);
  std::vector<T> x((size_t)(H.n > H.m ? H.n : H.m), (T)1), y((size_t)H.m, (T)0), yhat((size_t)H.m, (T)0);
  std::printf("---------BEFORE FIRST FILL---------\n");
  hspmv_matrix *A = nullptr;
  CHECK(hspmv_create_csr(&A, H.m, H.n, H.nnz, H.rp, H.ci, H.val, H.dtype, 0, 0));
  CHECK(hspmv_set_x(A, x.data()));
  test_spmv<T>(H, x.data(), yhat.data());
  std::printf("-----------After test_spmv--------\n");
  timed_runs(A, o.algo, N, /*warmup=*/0); // the reference CUDA driver has no warm-up (Q7)
  std::printf("TEAST\n");
  CHECK(hspmv_get_y(A, y.data(), 1));
  std::printf("Number Wrong: %d \n", count_wrong<T>(y.data(), yhat.data(), nullptr, H.m));
  hspmv_destroy(A);
  return 0;
}

int main(int argc, char **argv) {
  Options o = parse(argc, argv);
  if (o.pos.size() < 2) {
    std::printf("./spmv.exe inputfie.csr num_runs\n");
    return 0;
  }
  HostCsr H;
  H.dtype = o.fp32 ? HSPMV_F32 : HSPMV_F64;
  std::printf("Before Read\n");
  CHECK(read_matrix(o, o.pos[0], o.base < 0 ? 0 : o.base, H));
  std::printf("After Read\n");
  const int N = std::atoi(o.pos[1]);
  int rc = o.fp32 ? run<float>(o, H, N) : run<double>(o, H, N);
  hspmv_free_host(H.rp);
  hspmv_free_host(H.ci);
  hspmv_free_host(H.val);
  return rc;
}
