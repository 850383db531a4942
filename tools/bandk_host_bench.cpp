// Host-side pieces of the Band-k build, timed without a GPU (profiles/r2_host_bandk_cpu_container.txt):
//   * the RCM of a stand-in for config 5's level-1 coarse graph (gx x gy x gz vertices, 27-point neighbourhood with
//     multiplicities; 25 x 400 x 400 = 4 M vertices, 105 M entries), serial queue walk vs the level-synchronous walk,
//     and whether the two permutations are identical;
//   * the greedy grouping of 64 M rows (27-point row lengths), literal loop vs forward scan.
// Includes bandk.cpp itself to reach its file-local functions.
//   g++ -O3 -std=c++17 -fopenmp tools/bandk_host_bench.cpp -o /tmp/bandk_host_bench && /tmp/bandk_host_bench 25 400 400
#include "../heterogeneous-spmv_b200/csrc/bandk.cpp"

using namespace hspmv;

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
  const int gx = argc > 1 ? std::atoi(argv[1]) : 25, gy = argc > 2 ? std::atoi(argv[2]) : 400, gz = argc > 3 ? std::atoi(argv[3]) : 400;
  const size_t cutoff = argc > 4 ? (size_t)std::atoll(argv[4]) : 2048;
  Coarse g;
  g.n = (int64_t)gx * gy * gz;
  g.rp.assign((size_t)g.n + 1, 0);
  for (int z = 0; z < gz; z++)
    for (int y = 0; y < gy; y++)
      for (int x = 0; x < gx; x++) {
        const u32 v = (u32)(((int64_t)z * gy + y) * gx + x);
        for (int dz = -1; dz <= 1; dz++)
          for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
              const int X = x + dx, Y = y + dy, Z = z + dz;
              if (X < 0 || Y < 0 || Z < 0 || X >= gx || Y >= gy || Z >= gz) continue;
              g.adj.push_back((u32)(((int64_t)Z * gy + Y) * gx + X));
              g.mult.push_back(dx == 0 ? 46u + (v % 3) : 3u); // ties inside the x-line neighbours, as in the real coarse graph
            }
        g.rp[(size_t)v + 1] = (u32)g.adj.size();
      }
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  std::printf("coarse graph: %lld vertices, %zu entries; %d threads, frontier cutoff %zu\n", (long long)g.n, g.adj.size(), nt, cutoff);
  uvec a, b, c, d;
  for (int rep = 0; rep < 3; rep++) {
    double t0 = now();
    rcm_order(g, a, b);
    double t1 = now();
    rcm_order_parallel(g, nt, cutoff, c, d);
    double t2 = now();
    std::printf("RCM serial walk %.3f s   level-synchronous %.3f s   identical permutation: %s\n", t1 - t0, t2 - t1,
                (a == c && b == d) ? "yes" : "NO");
  }
  { // where the level-synchronous walk spends its time (first component only: the stand-in is connected)
    RcmPar s(g, nt, cutoff);
    uvec q((size_t)g.n);
    double t0 = now();
    const int root = s.peripheral(0);
    double t1 = now();
    s.number_component(root, 0, q);
    double t2 = now();
    std::printf("level-synchronous walk: pseudo-peripheral search (%zu levels) %.3f s, numbering %.3f s\n", s.lvl_ptr.size() - 1, t1 - t0, t2 - t1);
  }
  for (int k = 2; k <= 3; ++k) { // everything the device builder leaves to the host after the level-1 coarse graph: bandk_finish
    CoarseGraph level1;
    level1.n = g.n;
    level1.rp = g.rp;
    level1.adj = g.adj;
    level1.mult = g.mult;
    uvec map1((size_t)g.n + 1), perm, sr, ssr;
    for (int64_t i = 0; i <= g.n; ++i) map1[(size_t)i] = (u32)(16 * i); // 16 rows per level-1 group
    const int sizes[2] = {16, 16};
    double t0 = now();
    int rc = bandk_finish(16 * g.n, k, sizes, level1, map1, perm, sr, ssr);
    std::printf("bandk_finish k=%d on %lld rows: %.3f s (rc %d)  [HSPMV_BANDK_TRACE=1 prints the phases]\n", k, (long long)(16 * g.n), now() - t0, rc);
  }
  const int64_t n = argc > 5 ? std::atoll(argv[5]) : 64000000;
  uvec rp((size_t)n + 1);
  rp[0] = 0;
  for (int64_t i = 0; i < n; i++) rp[(size_t)i + 1] = rp[(size_t)i] + ((i % 400 == 0 || i % 400 == 399) ? 18 : 27);
  for (int rep = 0; rep < 3; rep++) {
    uvec s1, s2, grp;
    double t0 = now();
    group_rows_literal(rp.data(), n, 429, s1, grp);
    double t1 = now();
    group_starts(rp.data(), n, 429, s2);
    double t2 = now();
    std::printf("grouping of %lld rows: literal loop %.3f s   forward scan %.3f s   identical starts: %s (%zu groups)\n", (long long)n, t1 - t0,
                t2 - t1, s1 == s2 ? "yes" : "NO", s1.size());
  }
  return 0;
}
