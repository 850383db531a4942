// Drop-in for the offline converters of reformat-csr-to-csr3/ (no GPU needed):
//   reformat-auto <in.mtx.rcm.csr> <out.csr3>          spmv-auto.cpp:130-202 -> `.csr3` (k = 3, Volta
//                                                      heuristic sizes), byte-compatible writer :30-65
//   reformat <in> <out> [size0]                        spmv.cpp:130-197 -> Band-k reordered plain CSR (k = 2)
// 0-based reader (reformat-csr-to-csr3/spmv.cpp:95,102), fp32 values like the reference; --fp64 keeps doubles.
#include "common.h"

#include <cmath>

int main(int argc, char **argv) {
  Options o = parse(argc, argv);
  o.fp32 = true;
  for (int i = 1; i < argc; ++i)
    if (!std::strcmp(argv[i], "--fp64")) o.fp32 = false;
  const std::string self = argv[0];
  const bool to_csr3 = self.find("reformat-auto") != std::string::npos;
  if (o.pos.size() < 2) {
    std::printf("Syntax: %s inputfile outputfile\n", argv[0]);
    return 0;
  }
  HostCsr H;
  H.dtype = o.fp32 ? HSPMV_F32 : HSPMV_F64;
  CHECK(read_matrix(o, o.pos[0], o.base < 0 ? 0 : o.base, H));
  int sizes[2] = {0, 0};
  int k = to_csr3 ? 3 : 2;
  if (to_csr3) {
    CHECK(hspmv_auto_params(H.nnz, H.m, /*Volta*/ 0, sizes, nullptr, nullptr));
    std::printf("using ssrs %d, srs %d\n", sizes[0], sizes[1]);
    std::printf("SpMV\nHAND\n%d\n%d%d\n", k, sizes[0], sizes[1]); // spmv-auto.cpp:175-178 prints them unseparated
  } else if (o.pos.size() > 2) {
    sizes[0] = std::atoi(o.pos[2]);
  } else { // reformat-csr-to-csr3/spmv.cpp:154-170: super-row size from the density
    const double d = (double)H.nnz / (double)H.m;
    sizes[0] = (int)std::floor(10.14618 - 1.5 * std::log(d) + 0.5);
    if (sizes[0] < 1) sizes[0] = 1;
  }
  std::printf("Read in matrix and config file.\n");
  std::vector<uint32_t> perm((size_t)H.m), sr((size_t)H.m + 1), ssr((size_t)H.m + 1);
  int64_t n_sr = 0, n_ssr = 0;
  auto tic = std::chrono::steady_clock::now();
  CHECK(hspmv_bandk_host(H.m, H.nnz, (const uint32_t *)H.rp, (const uint32_t *)H.ci, k, sizes, perm.data(), sr.data(), &n_sr,
                         ssr.data(), &n_ssr));
  std::vector<uint32_t> rp2((size_t)H.m + 1), ci2((size_t)H.nnz);
  std::vector<unsigned char> v2((size_t)H.nnz * (o.fp32 ? 4 : 8));
  CHECK(hspmv_reorder_host(H.m, H.nnz, perm.data(), (const uint32_t *)H.rp, (const uint32_t *)H.ci, H.val, H.dtype, rp2.data(),
                           ci2.data(), v2.data()));
  std::chrono::duration<float> toc = std::chrono::steady_clock::now() - tic;
  std::printf("%s reordered in %g seconds.\n", o.pos[0], (double)toc.count());
  std::printf("In CSR-k format.\n");
  if (to_csr3)
    CHECK(hspmv_write_csr3_text(o.pos[1], H.dtype, n_ssr, n_sr, H.m, H.n, H.nnz, ssr.data(), sr.data(), rp2.data(), ci2.data(), v2.data()));
  else
    CHECK(hspmv_write_csr_text(o.pos[1], 0, H.dtype, H.m, H.n, H.nnz, (const int32_t *)rp2.data(), (const int32_t *)ci2.data(), v2.data()));
  hspmv_free_host(H.rp);
  hspmv_free_host(H.ci);
  hspmv_free_host(H.val);
  return 0;
}
