// Shared pieces of the drop-in drivers: the reference's CLI conventions and stdout protocol
// (`TimeMin: %lg`, `TimeMax: %lg`, `TimeAvg: %lg` in seconds, then `Number Wrong: %d `;
// run_scripts/run_cuda.py:110-122 slices exactly those prefixes).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hspmv.h"

struct Options {
  bool fp32 = false;   // --fp32: the reference's float arithmetic; default is fp64 (north_star)
  int base = -1;       // --base 0|1 overrides the driver's native index base
  int algo = HSPMV_ALGO_AUTO;
  int k = 0;           // --k 2|3
  bool cache = false;  // --cache: keep / use the binary side file <matrix>.hspmvbin
  int gpus = 1;        // --gpus N: row-shard the matrix over devices 0..N-1 (hspmv_mg_*, one host thread)
  std::vector<const char *> pos;
};

inline Options parse(int argc, char **argv) {
  Options o;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--fp32")) o.fp32 = true;
    else if (!std::strcmp(argv[i], "--fp64")) o.fp32 = false;
    else if (!std::strcmp(argv[i], "--cache")) o.cache = true;
    else if (!std::strcmp(argv[i], "--base") && i + 1 < argc) o.base = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) o.gpus = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--k") && i + 1 < argc) o.k = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--algo") && i + 1 < argc) {
      const char *a = argv[++i];
      o.algo = !std::strcmp(a, "vector") ? HSPMV_ALGO_CSR_VECTOR : !std::strcmp(a, "stream") ? HSPMV_ALGO_CSR_STREAM
             : !std::strcmp(a, "merge") ? HSPMV_ALGO_CSR_MERGE : !std::strcmp(a, "csrk") ? HSPMV_ALGO_CSRK : HSPMV_ALGO_AUTO;
    } else o.pos.push_back(argv[i]);
  }
  return o;
}

#define CHECK(call)                                                            \
  do {                                                                         \
    if (call) {                                                                \
      std::printf("Error: %s\n", hspmv_last_error());                          \
      std::exit(0); /* the reference drivers always exit with status 0 */      \
    }                                                                          \
  } while (0)

struct HostCsr {
  int64_t m = 0, n = 0, nnz = 0;
  int32_t *rp = nullptr, *ci = nullptr;
  void *val = nullptr;
  int dtype = HSPMV_F64;
  double value(int64_t i) const { return dtype == HSPMV_F64 ? ((double *)val)[i] : (double)((float *)val)[i]; }
};

// serial reference product on the host in the matrix's own precision -- the drivers' self-check
// (test_spmv, spmv-csr/spmv.c:68-90)
template <typename T> inline void test_spmv(const HostCsr &A, const T *x, T *y) {
  const T *v = (const T *)A.val;
  for (int64_t r = 0; r < A.m; ++r) {
    T acc = 0;
    for (int64_t j = A.rp[r]; j < A.rp[r + 1]; ++j) acc += v[j] * x[A.ci[j]];
    y[r] = acc;
  }
}

// text reader of the drivers, optionally through the binary side file
inline int read_matrix(const Options &o, const char *path, int base, HostCsr &H) {
  if (!o.cache) return hspmv_read_csr_text(path, base, H.dtype, &H.m, &H.n, &H.nnz, &H.rp, &H.ci, &H.val);
  int hit = 0;
  int rc = hspmv_read_csr_cached(path, base, H.dtype, &H.m, &H.n, &H.nnz, &H.rp, &H.ci, &H.val, &hit);
  if (!rc) std::printf("Matrix from %s\n", hit ? "the binary side file" : "text (side file written)");
  return rc;
}

struct Times {
  float mn = 9999.0f, mx = 0.0f, sum = 0.0f;
  int n = 0;
  void add(float t) {
    sum += t;
    if (t < mn) mn = t;
    if (t > mx) mx = t;
    ++n;
  }
  void print() const {
    std::printf("TimeMin: %lg\n", (double)mn);
    std::printf("TimeMax: %lg\n", (double)mx);
    std::printf("TimeAvg: %lg\n", (double)(sum / (n ? n : 1)));
  }
};

// N host-timed launch+sync iterations like the reference loops (cuda-spmv-csr/spmv.cu:240-260),
// plus the device-event times the host timer cannot see at this kernel length
inline void timed_runs(hspmv_matrix *A, int algo, int N, int warmup) {
  for (int i = 0; i < warmup; ++i) CHECK(hspmv_spmv(A, algo, nullptr, nullptr, nullptr));
  std::vector<float> sink(1);
  Times host;
  for (int i = 0; i < N; ++i) {
    auto tic = std::chrono::steady_clock::now();
    CHECK(hspmv_spmv(A, algo, nullptr, nullptr, nullptr));
    CHECK(hspmv_sync(A));
    std::chrono::duration<float> toc = std::chrono::steady_clock::now() - tic;
    host.add(toc.count());
  }
  host.print();
  if (N > 0) {
    std::vector<float> ms((size_t)N);
    CHECK(hspmv_spmv_timed(A, algo, 0, N, 0, ms.data()));
    float mn = 1e30f, sum = 0;
    for (float t : ms) { mn = t < mn ? t : mn; sum += t; }
    std::printf("DeviceTimeMin: %lg\nDeviceTimeAvg: %lg\n", (double)mn * 1e-3, (double)(sum / N) * 1e-3);
    std::printf("Kernel: %s\n", hspmv_last_kernel(A));
  }
}

template <typename T> inline int count_wrong(const T *y, const T *yref, const unsigned int *perm, int64_t m) {
  int wrong = 0;
  for (int64_t i = 0; i < m; ++i) {
    const double d = (double)y[i] - (double)yref[perm ? perm[i] : i];
    if (d > .01 || d < -.01) ++wrong;
  }
  return wrong;
}

#define CHECK_MG(call)                                                         \
  do {                                                                         \
    if (call) {                                                                \
      std::printf("Error: %s\n", hspmv_mg_last_error());                       \
      std::exit(0);                                                            \
    }                                                                          \
  } while (0)

// --gpus N: the same protocol on N devices of the box through the multi-GPU entry points of the C ABI.
// x and y cross hspmv_mg_* in the ORIGINAL numbering, so the self-check needs no permutation.
// k = 1: plain CSR; k = 2 / 3: Band-k on device 0, shards cut at super-row / super-super-row boundaries.
template <typename T>
inline int run_multi_gpu(const Options &o, const HostCsr &H, int N, int warmup, int k, const int *sizes, int index_base) {
  std::vector<int> dev((size_t)o.gpus);
  for (int i = 0; i < o.gpus; ++i) dev[(size_t)i] = i;
  hspmv_mg *M = nullptr;
  CHECK_MG(hspmv_mg_create_csr(&M, dev.data(), o.gpus, H.m, H.n, H.nnz, H.rp, H.ci, H.val, H.dtype, index_base, k, sizes));
  int repl = 0;
  CHECK_MG(hspmv_mg_info(M, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &repl));
  std::vector<int64_t> row_lo((size_t)o.gpus + 1), halo((size_t)o.gpus), nz((size_t)o.gpus);
  CHECK_MG(hspmv_mg_partition_query(M, nullptr, row_lo.data(), nullptr, nullptr, nz.data(), halo.data()));
  std::printf("Multi-GPU: %d devices, %s\n", o.gpus, repl ? "x replicated, y row-sharded" : "nnz-balanced row blocks + x halo exchange over NVLink");
  for (int d = 0; d < o.gpus; ++d)
    std::printf("  device %d: rows [%lld,%lld) nnz %lld halo %lld\n", d, (long long)row_lo[(size_t)d], (long long)row_lo[(size_t)d + 1],
                (long long)nz[(size_t)d], (long long)halo[(size_t)d]);
  std::vector<T> x((size_t)(H.n > H.m ? H.n : H.m), (T)1), y((size_t)H.m, (T)0), yser((size_t)H.m, (T)0);
  CHECK_MG(hspmv_mg_set_x(M, x.data()));
  for (int i = 0; i < warmup; ++i) CHECK_MG(hspmv_mg_spmv(M));
  CHECK_MG(hspmv_mg_sync(M));
  Times host;
  for (int i = 0; i < N; ++i) {
    auto tic = std::chrono::steady_clock::now();
    CHECK_MG(hspmv_mg_spmv(M));
    CHECK_MG(hspmv_mg_sync(M));
    std::chrono::duration<float> toc = std::chrono::steady_clock::now() - tic;
    host.add(toc.count());
  }
  host.print();
  if (N > 0) {
    std::vector<float> ms((size_t)N);
    CHECK_MG(hspmv_mg_spmv_timed(M, 0, N, ms.data()));
    float mn = 1e30f, sum = 0;
    for (float t : ms) { mn = t < mn ? t : mn; sum += t; }
    std::printf("DeviceTimeMin: %lg\nDeviceTimeAvg: %lg\n", (double)mn * 1e-3, (double)(sum / N) * 1e-3);
    std::printf("Kernel: %s\n", hspmv_last_kernel((hspmv_matrix *)hspmv_mg_shard(M, 0)));
  }
  HostCsr Z = H; // the self-check product on 0-based arrays
  std::vector<int32_t> rp0, ci0;
  if (index_base) {
    rp0.assign(H.rp, H.rp + H.m + 1);
    ci0.assign(H.ci, H.ci + H.nnz);
    for (auto &t : rp0) t -= index_base;
    for (auto &t : ci0) t -= index_base;
    Z.rp = rp0.data();
    Z.ci = ci0.data();
  }
  test_spmv<T>(Z, x.data(), yser.data());
  CHECK_MG(hspmv_mg_get_y(M, y.data()));
  std::printf("Number Wrong: %d \n", count_wrong<T>(y.data(), yser.data(), nullptr, H.m));
  hspmv_mg_destroy(M);
  return 0;
}
