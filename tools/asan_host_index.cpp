// AddressSanitizer / UBSan pass over the host index work (Band-k with both RCM walks, STS level sets, bins, blocks, tiles,
// partition, pipeline plan) on 200 random symmetric matrices:
//   cd heterogeneous-spmv_b200/csrc && g++ -O1 -g -std=c++17 -fopenmp -fsanitize=address,undefined ../../tools/asan_host_index.cpp bandk.cpp sts.cpp partition.cpp -o /tmp/asan_host
//   HSPMV_RCM=parallel HSPMV_RCM_CUTOFF=3 ASAN_OPTIONS=detect_leaks=0 /tmp/asan_host   # and once without the variables
#include "../heterogeneous-spmv_b200/csrc/host_index.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
using namespace hspmv;
int main(int argc, char **argv) {
  std::mt19937 rng(5);
  const int trials = argc > 1 ? std::atoi(argv[1]) : 200;
  for (int trial = 0; trial < trials; ++trial) {
    int n = 2 + rng() % 3000;
    std::vector<std::set<int>> adj(n);
    int e = n * (1 + rng() % 6);
    for (int i = 0; i < e; ++i) { int a = rng() % n, b = rng() % n; adj[a].insert(b); adj[b].insert(a); }
    for (int i = 0; i < n; ++i) adj[i].insert(i);
    if (trial % 7 == 0) for (int i = 0; i + 1 < n; ++i) { adj[i].insert(i + 1); adj[i + 1].insert(i); }
    std::vector<int32_t> rp(1, 0), ci;
    for (int i = 0; i < n; ++i) { for (int c : adj[i]) ci.push_back(c); rp.push_back((int32_t)ci.size()); }
    uvec perm; std::vector<int32_t> pack;
    int rc = sts_levels(n, rp.data(), ci.data(), perm, pack);
    if (rc || (int)perm.size() != n || pack.back() != n) { printf("sts fail %d\n", rc); return 1; }
    std::vector<int32_t> rows; int64_t bp[NUM_BINS + 1];
    bin_rows(n, rp.data(), rows, bp);
    int64_t st[10]; row_stats(n, rp.data(), st);
    row_len_percentile(n, rp.data(), 90);
    std::vector<int32_t> blk; block_rows(n, rp.data(), 64, 16, blk);
    std::vector<int32_t> tr; tile_rows(n, rp.data(), 37, tr);
    int64_t lo[9]; partition_rows(n, rp.data(), 8, nullptr, 0, lo);
    uvec hc((size_t)n), hp, hi;
    for (int i = 0; i < n; ++i) hc[(size_t)i] = (trial % 5 == 0 && i == n / 2) ? (u32)(20 * n) : (u32)(rng() % 40);
    hub_order(n, hc.data(), hp, hi);
    for (int i = 0; i < n; ++i) if (hi[hp[(size_t)i]] != (u32)i) { printf("hub_order fail\n"); return 1; }
    uvec p2, sr, ssr; int sizes[2] = {1 + (int)(rng() % 9), 1 + (int)(rng() % 5)};
    int k = 2 + trial % 2;
    rc = bandk_build(n, (int64_t)ci.size(), (const u32 *)rp.data(), (const u32 *)ci.data(), k, sizes, p2, sr, ssr);
    if (rc) { printf("bandk fail %d\n", rc); return 1; }
  }
  for (int64_t m : {4200000LL, 64000000LL, 2147483000LL}) {
    std::vector<int64_t> rl, xl; std::vector<int> need;
    pipe_chunks(m, 8, rl);
    std::vector<int64_t> cm(rl.size() - 1);
    for (size_t c = 0; c + 1 < rl.size(); ++c) cm[c] = std::min<int64_t>(m - 1, rl[c + 1] + 1000);
    pipe_pieces(m, cm, xl, need);
  }
  { // called from inside a caller's parallel region (the inner teams then have ONE thread): same answers
    const int64_t m = 200000;
    std::vector<int32_t> rp((size_t)m + 1, 0);
    uvec cnt((size_t)m);
    for (int64_t i = 0; i < m; ++i) { rp[(size_t)i + 1] = rp[(size_t)i] + (int32_t)(rng() % 70); cnt[(size_t)i] = (u32)(rng() % 1000); }
    std::vector<int32_t> rows0; int64_t bp0[NUM_BINS + 1]; bin_rows(m, rp.data(), rows0, bp0);
    uvec p0, i0; hub_order(m, cnt.data(), p0, i0);
    int bad = 0;
#pragma omp parallel num_threads(2) reduction(+ : bad)
    {
      std::vector<int32_t> rows1; int64_t bp1[NUM_BINS + 1]; bin_rows(m, rp.data(), rows1, bp1);
      uvec p1, i1; hub_order(m, cnt.data(), p1, i1);
      if (rows1 != rows0 || p1 != p0 || i1 != i0) bad = 1;
      for (int b = 0; b <= NUM_BINS; ++b) if (bp1[b] != bp0[b]) bad = 1;
    }
    if (bad) { printf("nested call fail\n"); return 1; }
  }
  printf("ok\n");
}
