// AddressSanitizer / UBSan pass over the text readers (csrc/io.cpp) on valid, truncated, corrupted and absurd-header files,
// small (single-piece parse) and large (multi-piece parse):
//   cd heterogeneous-spmv_b200/csrc && g++ -O1 -g -std=c++17 -fopenmp -fsanitize=address,undefined ../../tools/asan_io.cpp io.cpp -o /tmp/asan_io
//   ASAN_OPTIONS=detect_leaks=0 /tmp/asan_io
#include "../heterogeneous-spmv_b200/csrc/host_index.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <algorithm>
#include <string>
#include <vector>
using namespace hspmv;
static void put(const std::string &path, const std::string &s) { FILE *f = fopen(path.c_str(), "wb"); fwrite(s.data(), 1, s.size(), f); fclose(f); }
int main() {
  std::mt19937 rng(9);
  const std::string path = "/tmp/asan_io_case.csr";
  int ok = 0, rejected = 0;
  for (int trial = 0; trial < 600; ++trial) {
    // a valid small file ...
    int m = (trial % 10 == 9) ? 20000 + (int)(rng() % 5000) : 1 + (int)(rng() % 40), base = rng() % 2;   // large ones take the multi-piece parse
    std::string s;
    std::vector<int> rp(1, 0), ci;
    for (int r = 0; r < m; ++r) { int len = rng() % 5; for (int j = 0; j < len; ++j) ci.push_back(rng() % m); rp.push_back((int)ci.size()); }
    s += std::to_string(m) + " " + std::to_string(m) + " " + std::to_string(ci.size()) + "\n";
    for (int v : rp) s += std::to_string(v + base) + " ";
    s += "\n";
    for (int v : ci) s += std::to_string(v + base) + " ";
    s += "\n";
    for (size_t j = 0; j < ci.size(); ++j) { char b[64]; snprintf(b, sizeof b, "%f ", (double)((int)(rng() % 2000) - 1000) / 64.0); s += b; }
    s += "\n";
    // ... mutilated in every third trial: truncation, garbage bytes, huge counts, negative numbers, no trailing newline
    const int kind = trial % 6;
    if (kind == 1) s = s.substr(0, rng() % (s.size() + 1));
    if (kind == 2) for (int t = 0; t < 3 && !s.empty(); ++t) s[rng() % s.size()] = (char)(rng() % 256);
    if (kind == 3) s = "2000000000 2000000000 2000000000\n" + s;
    if (kind == 4) s = "-5 3 -1\n" + s;
    if (kind == 5 && !s.empty()) s.pop_back();
    put(path, s);
    for (int dtype : {HSPMV_F32, HSPMV_F64}) {
      int64_t M = 0, N = 0, NNZ = 0; int32_t *a = nullptr, *b = nullptr; void *v = nullptr; std::string err;
      int rc = read_csr_text(path.c_str(), base, dtype, &M, &N, &NNZ, &a, &b, &v, err);
      if (rc == 0) { ++ok; free(a); free(b); free(v); } else ++rejected;
    }
  }
  // .csr3: header "nSSR nSR M N NNZ", then ssr_ptr, sr_ptr, row_ptr, col, val on one line (reformat-csr-to-csr3/spmv-auto.cpp:38-62)
  int ok3 = 0, rejected3 = 0;
  for (int trial = 0; trial < 300; ++trial) {
    int m = (trial % 10 == 9) ? 30000 : 2 + (int)(rng() % 60);
    std::vector<unsigned> rp(1, 0), ci, sr(1, 0), ssr(1, 0);
    for (int r = 0; r < m; ++r) { int len = rng() % 4; for (int j = 0; j < len; ++j) ci.push_back(rng() % m); rp.push_back((unsigned)ci.size()); }
    for (unsigned r = 0; r < (unsigned)m;) { r = std::min<unsigned>((unsigned)m, r + 1 + rng() % 5); sr.push_back(r); }
    for (unsigned q = 0; q + 1 < sr.size();) { q = std::min<unsigned>((unsigned)sr.size() - 1, q + 1 + rng() % 3); ssr.push_back(q); }
    std::string s = std::to_string(ssr.size() - 1) + " " + std::to_string(sr.size() - 1) + " " + std::to_string(m) + " " + std::to_string(m) + " " +
                    std::to_string(ci.size()) + " \n";
    for (unsigned v : ssr) s += std::to_string(v) + " ";
    for (unsigned v : sr) s += std::to_string(v) + " ";
    for (unsigned v : rp) s += std::to_string(v) + " ";
    for (unsigned v : ci) s += std::to_string(v) + " ";
    for (size_t j = 0; j < ci.size(); ++j) { char b[64]; snprintf(b, sizeof b, "%.6f ", (double)((int)(rng() % 2000) - 1000) / 64.0); s += b; }
    const int kind = trial % 5;
    if (kind == 1) s = s.substr(0, rng() % (s.size() + 1));
    if (kind == 2) for (int t = 0; t < 3 && !s.empty(); ++t) s[rng() % s.size()] = (char)(rng() % 256);
    if (kind == 3) s = "4000000000 4000000000 4000000000 4000000000 4000000000 \n" + s;
    put(path, s);
    for (int dtype : {HSPMV_F32, HSPMV_F64}) {
      int64_t a1 = 0, a2 = 0, M = 0, N = 0, NNZ = 0; u32 *p1 = nullptr, *p2 = nullptr, *p3 = nullptr, *p4 = nullptr; void *v = nullptr; std::string err;
      int rc = read_csr3_text(path.c_str(), dtype, &a1, &a2, &M, &N, &NNZ, &p1, &p2, &p3, &p4, &v, err);
      if (rc == 0) { ++ok3; free(p1); free(p2); free(p3); free(p4); free(v); } else ++rejected3;
    }
  }
  printf("csr ok %d rejected %d; csr3 ok %d rejected %d\n", ok, rejected, ok3, rejected3);
  return 0;
}
