// Drop-in for the CSR-k drivers:
//   spmv <file> <num_runs> <size0> [<size1>]     cuda-spmv-csrk/cuda/spmv.cu:111-266, spmv-csrk/spmv.cpp:97-244
//   spmv-auto-volta|ampere|b200 <file> <num_runs>   cuda-spmv-csrk/cuda/spmv-auto-*.cu:111-287
// (the auto flavour is selected by the executable's name or by --auto volta|ampere|mi100|b200).
// 1-based reader (spmv-csrk/spmv.cpp:60,67).  argv[3] sizes level 1 (rows per super-row) and
// argv[4] level 2, whatever the reference's usage string calls them (quirk Q9).  k = 2 with one
// size, 3 with two (override: --k).  The self-check compares y[i] with y_csrser[perm[i]] as the
// CPU driver does (spmv-csrk/spmv.cpp:202-211): the reordered matrix is what runs on the device.
#include "common.h"

static bool o_tune = false; // --tune, and the default of spmv-auto-b200 (--no-tune switches it off)

template <typename T> static int run(const Options &o, HostCsr &H, int N, int k, int *sizes, bool print_perm) {
  if (o.gpus > 1) return run_multi_gpu<T>(o, H, N, /*warmup=*/5, k, sizes, 0); // (read_matrix already removed the 1-base)
  std::vector<T> x((size_t)H.m, (T)1), y((size_t)H.m, (T)0), yser((size_t)H.m, (T)0);
  hspmv_matrix *A = nullptr;
  CHECK(hspmv_create_csr(&A, H.m, H.n, H.nnz, H.rp, H.ci, H.val, H.dtype, 0, 1));
  auto tic = std::chrono::steady_clock::now();
  CHECK(hspmv_build_csrk(A, k, sizes));
  std::chrono::duration<float> toc = std::chrono::steady_clock::now() - tic;
  std::printf("In CSR-k format.\n");
  int64_t n_sr = 0, n_ssr = 0;
  hspmv_matrix_info(A, nullptr, nullptr, nullptr, nullptr, nullptr, &n_sr, &n_ssr);
  std::printf("Band-k: %lld super-rows, %lld super-super-rows, built in %g seconds.\n", (long long)n_sr, (long long)n_ssr, (double)toc.count());
  CHECK(hspmv_set_x(A, x.data()));
  if (o_tune) { // spmv-auto-b200: measure instead of trusting a formula
    float tuned = 0;
    CHECK(hspmv_autotune(A, HSPMV_ALGO_CSRK, 1, &tuned));
    std::printf("Tuned: %s (%g ms)\n", hspmv_last_kernel(A), (double)tuned);
  }
  std::printf("Vectors set, running test.\n");
  timed_runs(A, HSPMV_ALGO_CSRK, N, /*warmup=*/5); // 5 warm-ups as spmv-csrk/spmv.cpp:169-171
  test_spmv<T>(H, x.data(), yser.data());
  std::vector<unsigned int> perm((size_t)H.m);
  CHECK(hspmv_csrk_get(A, perm.data(), nullptr, nullptr));
  if (print_perm) { // spmv-csrk/spmv.cpp:199-201
    for (int64_t i = 0; i < H.m; ++i) std::printf("%d ", (int)perm[(size_t)i]);
    std::printf("\n");
  }
  CHECK(hspmv_get_y(A, y.data(), 0));
  std::printf("Number Wrong: %d \n", count_wrong<T>(y.data(), yser.data(), perm.data(), H.m));
  hspmv_destroy(A);
  return 0;
}

int main(int argc, char **argv) {
  Options o;
  int arch = -1;
  bool print_perm = false, no_tune = false;
  {
    std::vector<char *> keep;
    keep.push_back(argv[0]);
    for (int i = 1; i < argc; ++i) {
      if (!std::strcmp(argv[i], "--auto") && i + 1 < argc) {
        const char *a = argv[++i];
        arch = !std::strcmp(a, "volta") ? 0 : !std::strcmp(a, "ampere") ? 1 : !std::strcmp(a, "mi100") ? 2 : 3;
      } else if (!std::strcmp(argv[i], "--print-permutation")) print_perm = true;
      else if (!std::strcmp(argv[i], "--tune")) o_tune = true;
      else if (!std::strcmp(argv[i], "--no-tune")) no_tune = true;
      else keep.push_back(argv[i]);
    }
    o = parse((int)keep.size(), keep.data());
  }
  const std::string self = argv[0];
  if (arch < 0) {
    if (self.find("auto-volta") != std::string::npos) arch = 0;
    else if (self.find("auto-ampere") != std::string::npos) arch = 1;
    else if (self.find("auto-mi100") != std::string::npos) arch = 2;
    else if (self.find("auto-b200") != std::string::npos) arch = 3;
  }
  const bool is_auto = arch >= 0;
  if (arch == 3 && !no_tune) o_tune = true;
  if (o.pos.size() < (is_auto ? 2u : 3u)) {
    if (is_auto) std::printf("Syntax: %s inputfile num_runs\n", argv[0]);
    else std::printf("Syntax: %s inputfile num_runs super_super_row_size super_row_size\n", argv[0]);
    return 0;
  }
  HostCsr H;
  H.dtype = o.fp32 ? HSPMV_F32 : HSPMV_F64;
  CHECK(read_matrix(o, o.pos[0], o.base < 0 ? 1 : o.base, H));
  std::printf("size %d %d %d \n", (int)H.m, (int)H.n, (int)H.nnz);
  const int N = std::atoi(o.pos[1]);
  int sizes[2] = {0, 0}, k;
  if (is_auto) {
    int vec = 0, bdy = 0;
    CHECK(hspmv_auto_params(H.nnz, H.m, arch, sizes, &vec, &bdy));
    k = o.k ? o.k : 3;
    std::printf("using ssrs %d, srs %d\n", sizes[0], sizes[1]);
    if (vec) std::printf("Block dim is %dx%dx%d\n", vec, 8, bdy); else std::printf("Block dim is %dx%d\n", 8, bdy);
  } else {
    sizes[0] = std::atoi(o.pos[2]);
    if (o.pos.size() > 3) sizes[1] = std::atoi(o.pos[3]);
    k = o.k ? o.k : (o.pos.size() > 3 ? 3 : 2);
    if (k == 3 && sizes[1] <= 0) sizes[1] = 8;
  }
  std::printf("SpMV\nHAND\n%d\n", k);
  std::printf("Read in matrix and config file.\n");
  int rc = o.fp32 ? run<float>(o, H, N, k, sizes, print_perm) : run<double>(o, H, N, k, sizes, print_perm);
  hspmv_free_host(H.rp);
  hspmv_free_host(H.ci);
  hspmv_free_host(H.val);
  return rc;
}
