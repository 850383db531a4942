// Integer index work that feeds the kernels and the multi-GPU layer: nnz-balanced row blocks,
// column ranges (halo discovery), row-length bins, merge-path tile table.  All of it must be
// bit-exact against oracle/partition_oracle.py, which restates the same rules in numpy
// (the reference has no multi-GPU or binning code to compare with -- SURVEY.md 8e).
#include "host_index.h"

#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace hspmv {

int partition_rows(int64_t m, const int32_t *rp, int n_parts, const u32 *align_ptr,
                   int64_t n_align, int64_t *row_lo) {
  if (m < 0 || !rp || n_parts < 1 || !row_lo) return HSPMV_EINVAL;
  const int64_t nnz = rp[m];
  row_lo[0] = 0;
  for (int p = 1; p < n_parts; ++p) {
    const int64_t target = (int64_t)(((__int128)nnz * p) / n_parts);
    // first row r with rp[r] >= target
    int64_t r = std::lower_bound(rp, rp + m + 1, (int32_t)target) - rp;
    if (r > m) r = m;
    if (align_ptr) { // snap up to the next group boundary (super-row or super-super-row start)
      const u32 *b = std::lower_bound(align_ptr, align_ptr + n_align + 1, (u32)r);
      r = (b == align_ptr + n_align + 1) ? m : (int64_t)*b;
    }
    row_lo[p] = std::max(r, row_lo[p - 1]);
  }
  row_lo[n_parts] = m;
  return HSPMV_OK;
}

int block_col_range(const int32_t *rp, const int32_t *ci, int64_t row_lo, int64_t row_hi,
                    int64_t *cmin, int64_t *cmax) {
  if (!rp || !ci || !cmin || !cmax || row_hi < row_lo) return HSPMV_EINVAL;
  int32_t lo = INT32_MAX, hi = -1;
  const int64_t e0 = rp[row_lo], e1 = rp[row_hi];
#pragma omp parallel for reduction(min : lo) reduction(max : hi) schedule(static)
  for (int64_t e = e0; e < e1; ++e) {
    lo = std::min(lo, ci[e]);
    hi = std::max(hi, ci[e]);
  }
  *cmin = (e1 > e0) ? lo : 0;
  *cmax = (e1 > e0) ? hi : -1;
  return HSPMV_OK;
}

// These three walk every row of matrices with up to 64 M rows (twice during a Band-k build), so they run on all host
// cores: per-thread counts over contiguous row ranges, combined in thread order (the results do not depend on the
// number of threads; bin_rows keeps the rows of a bin in ascending order).
void row_stats(int64_t m, const int32_t *rp, int64_t out[10]) {
  for (int i = 0; i < 10; ++i) out[i] = 0;
  int64_t mx = 0;
#pragma omp parallel if (m > (1 << 16))
  {
    int64_t loc[NUM_BINS] = {0}, lmx = 0;
#pragma omp for schedule(static) nowait
    for (int64_t r = 0; r < m; ++r) {
      const int64_t len = (int64_t)rp[r + 1] - rp[r];
      ++loc[bin_of_len(len)];
      lmx = std::max(lmx, len);
    }
#pragma omp critical
    {
      for (int b = 0; b < NUM_BINS; ++b) out[b] += loc[b];
      mx = std::max(mx, lmx);
    }
  }
  out[8] = mx;
  out[9] = out[0];
}

// smallest length L such that at least pct % of the rows have <= L entries
int64_t row_len_percentile(int64_t m, const int32_t *rp, int pct) {
  if (m <= 0) return 0;
  const int64_t cap = 1 << 16;
  std::vector<int64_t> hist((size_t)cap + 1, 0);
#pragma omp parallel if (m > (1 << 16))
  {
    std::vector<int64_t> loc((size_t)cap + 1, 0);
#pragma omp for schedule(static) nowait
    for (int64_t r = 0; r < m; ++r) {
      const int64_t len = (int64_t)rp[r + 1] - rp[r];
      ++loc[(size_t)(len < cap ? len : cap)];
    }
#pragma omp critical
    for (int64_t l = 0; l <= cap; ++l) hist[(size_t)l] += loc[(size_t)l];
  }
  const int64_t need = (m * pct + 99) / 100;
  int64_t seen = 0;
  for (int64_t l = 0; l <= cap; ++l) {
    seen += hist[(size_t)l];
    if (seen >= need) return l;
  }
  return cap;
}

void bin_rows(int64_t m, const int32_t *rp, std::vector<int32_t> &rows, int64_t bin_ptr[NUM_BINS + 1]) {
  int nt = 1;
#ifdef _OPENMP
  nt = m > (1 << 16) ? omp_get_max_threads() : 1;
#endif
  // cnt[t][b] = rows of bin b in thread t's contiguous row range
  std::vector<int64_t> cnt((size_t)nt * NUM_BINS, 0);
  // (a loop over the nt row ranges, not "one range per thread of the team": inside a caller's parallel region the team
  // may be smaller than requested, and every range must still be walked)
#pragma omp parallel for schedule(static, 1) num_threads(nt)
  for (int t = 0; t < nt; ++t) {
    const int64_t lo = m * t / nt, hi = m * (t + 1) / nt;
    int64_t *c = &cnt[(size_t)t * NUM_BINS];
    for (int64_t r = lo; r < hi; ++r) ++c[bin_of_len((int64_t)rp[r + 1] - rp[r])];
  }
  bin_ptr[0] = 0;
  for (int b = 0; b < NUM_BINS; ++b) {
    int64_t tot = 0;
    for (int t = 0; t < nt; ++t) tot += cnt[(size_t)t * NUM_BINS + b];
    bin_ptr[b + 1] = bin_ptr[b] + tot;
  }
  rows.resize((size_t)m);
#pragma omp parallel for schedule(static, 1) num_threads(nt)
  for (int t = 0; t < nt; ++t) {
    int64_t cur[NUM_BINS];
    for (int b = 0; b < NUM_BINS; ++b) {
      cur[b] = bin_ptr[b];
      for (int u = 0; u < t; ++u) cur[b] += cnt[(size_t)u * NUM_BINS + b];
    }
    const int64_t lo = m * t / nt, hi = m * (t + 1) / nt;
    for (int64_t r = lo; r < hi; ++r) rows[(size_t)cur[bin_of_len((int64_t)rp[r + 1] - rp[r])]++] = (int32_t)r;
  }
}

void block_rows(int64_t m, const int32_t *rp, int64_t cap, int64_t max_rows, std::vector<int32_t> &blk_row) {
  blk_row.clear();
  blk_row.push_back(0);
  int64_t start = 0;
  for (int64_t r = 0; r < m; ++r) {
    // would row r still fit into the block that starts at `start`?
    if (r > start && ((int64_t)rp[r + 1] - rp[start] > cap || r - start >= max_rows)) {
      blk_row.push_back((int32_t)r);
      start = r;
    }
  }
  if (m > 0) blk_row.push_back((int32_t)m);
}

// ---- plan of the pipelined host call (hspmv_spmv_host, abi.cu:spmv_host_pipelined) ----------------------------------
// Row chunks: y goes down (and the kernel runs) chunk by chunk.  Equal chunks of ~4 MB of y, at least 4 and at most 32
// of them (the plan every r2 measurement was taken with: 32 x 16 MB on config 5, 16 x 4 MB on the 64 MB shards of an
// 8-GPU run); boundaries are multiples of 4 rows (32 B-aligned copies).  A timeline model of the three legs
// (tools/pipe_model.py) says that chunk sizes ramped up at the start and down at the end buy nothing -- the y leg
// always trails the x leg by one regular chunk -- so the sizes stay equal; what does pay is where the x pieces are cut
// (pipe_pieces below).
void pipe_chunks(int64_t m, int64_t vs, std::vector<int64_t> &row_lo) {
  row_lo.assign(1, 0);
  if (m <= 0) {
    row_lo.push_back(0);
    return;
  }
  int64_t C = m * vs / ((int64_t)4 << 20);
  if (C < 4) C = 4;
  if (C > 32) C = 32;
  for (int64_t c = 1; c < C; ++c) {
    int64_t r = ((m * c / C) + 3) & ~(int64_t)3;
    if (r > m) r = m;
    if (r < row_lo.back()) r = row_lo.back();
    row_lo.push_back(r);
  }
  row_lo.push_back(m);
}

// x pieces, one per row chunk: piece c ends right behind the highest column chunk c references (chunk_cmax[c], -1 for a
// chunk without entries), so that for a banded matrix chunk c waits for piece c only -- with equal pieces it waited for
// piece c + 1, because a block of rows references a little more than its own slice of x.  Piece ends never decrease and
// the last piece ends at n; need[c] = the last piece chunk c has to wait for (pieces are uploaded in order).
void pipe_pieces(int64_t n, const std::vector<int64_t> &chunk_cmax, std::vector<int64_t> &x_lo, std::vector<int> &need) {
  const size_t C = chunk_cmax.size();
  x_lo.assign(C + 1, 0);
  for (size_t c = 0; c < C; ++c) {
    int64_t e = ((chunk_cmax[c] + 1) + 3) & ~(int64_t)3;
    if (e > n) e = n;
    if (e < x_lo[c]) e = x_lo[c];
    x_lo[c + 1] = e;
  }
  if (C) x_lo[C] = n;
  need.assign(C, 0);
  for (size_t c = 0; c < C; ++c) {
    int piece = 0;
    while ((size_t)piece + 1 < C && x_lo[(size_t)piece + 1] <= chunk_cmax[c]) ++piece;
    need[c] = piece;
  }
}

// Order of the columns for hspmv_relabel_hub_columns: by descending reference count, ascending column id inside a
// count (perm[new] = old, inv[old] = new).  A stable counting sort on key = max - count; on all cores when the
// per-thread histograms stay small (threads take contiguous column ranges, so concatenating their slots per key in
// thread order keeps the ids ascending), else on one.
void hub_order(int64_t n, const u32 *cnt, uvec &perm, uvec &inv) {
  perm.resize((size_t)(n > 0 ? n : 0));
  inv.resize((size_t)(n > 0 ? n : 0));
  if (n <= 0) return;
  u32 mx = 0;
#pragma omp parallel for schedule(static) reduction(max : mx) if (n > (1 << 16))
  for (int64_t c = 0; c < n; ++c) mx = std::max(mx, cnt[c]);
  int nt = 1;
#ifdef _OPENMP
  nt = n > (1 << 16) ? omp_get_max_threads() : 1;
#endif
  const int64_t keys = (int64_t)mx + 1;
  if (keys > 4 * n + 1024) { // a count range far larger than the number of columns: no histogram at all
    for (int64_t c = 0; c < n; ++c) perm[(size_t)c] = (u32)c;
    std::stable_sort(perm.begin(), perm.end(), [&](u32 a, u32 b) { return cnt[a] > cnt[b]; });
    for (int64_t i = 0; i < n; ++i) inv[perm[(size_t)i]] = (u32)i;
    return;
  }
  if (keys * nt > n) nt = 1; // histograms larger than the data: one core, one histogram
  std::vector<int64_t> slot((size_t)(keys * nt), 0); // slot[t * keys + k]: first count, then first output position
#pragma omp parallel for schedule(static, 1) num_threads(nt)
  for (int t = 0; t < nt; ++t) {
    int64_t *h = &slot[(size_t)t * (size_t)keys];
    const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
    for (int64_t c = lo; c < hi; ++c) ++h[mx - cnt[c]];
  }
  int64_t run = 0;
  for (int64_t k = 0; k < keys; ++k)
    for (int t = 0; t < nt; ++t) {
      int64_t &v = slot[(size_t)t * (size_t)keys + (size_t)k];
      const int64_t c = v;
      v = run;
      run += c;
    }
#pragma omp parallel for schedule(static, 1) num_threads(nt)
  for (int t = 0; t < nt; ++t) {
    int64_t *h = &slot[(size_t)t * (size_t)keys];
    const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
    for (int64_t c = lo; c < hi; ++c) perm[(size_t)h[mx - cnt[c]]++] = (u32)c;
  }
#pragma omp parallel for schedule(static) if (n > (1 << 16))
  for (int64_t i = 0; i < n; ++i) inv[perm[(size_t)i]] = (u32)i;
}

void tile_rows(int64_t m, const int32_t *rp, int64_t tile, std::vector<int32_t> &tile_row) {
  const int64_t nnz = rp[m];
  const int64_t n_tiles = (nnz + tile - 1) / tile;
  tile_row.resize((size_t)n_tiles + 1);
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < n_tiles; ++t) {
    const int64_t e = t * tile;
    // last row r with rp[r] <= e  (the row that owns nonzero e)
    const int32_t *u = std::upper_bound(rp, rp + m + 1, (int32_t)e);
    tile_row[(size_t)t] = (int32_t)((u - rp) - 1);
  }
  tile_row[(size_t)n_tiles] = (int32_t)m;
}

} // namespace hspmv
