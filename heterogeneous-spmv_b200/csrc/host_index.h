// Host-side index work shared by the C ABI (abi.cu) and the converters: Band-k builder,
// text I/O, partitioning, binning, size heuristics.  No CUDA in here.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/hspmv.h"

namespace hspmv {

typedef uint32_t u32;

// uvec whose resize() / sized constructor leave the new elements UNINITIALISED (value-constructing forms --
// v(n, 0), assign(n, 0) -- still write).  The index arrays of a Band-k build at config 5 are 256 MB (permutation) to
// 420 MB (coarse graph) each; std::vector would zero them on one core -- page faults included, 0.15-0.25 s apiece --
// before the parallel loop or the download that overwrites every element.  With this type the first touch happens in
// that parallel loop, or in first_touch() right before a download.
template <typename T> struct RawAlloc {
  typedef T value_type;
  RawAlloc() noexcept {}
  template <class U> RawAlloc(const RawAlloc<U> &) noexcept {}
  T *allocate(size_t n) { return static_cast<T *>(::operator new(n * sizeof(T))); }
  void deallocate(T *p, size_t) noexcept { ::operator delete(p); }
  template <class U> void construct(U *p) noexcept { ::new ((void *)p) U; } // default-init: nothing for u32
  template <class U, class... A> void construct(U *p, A &&...a) { ::new ((void *)p) U(static_cast<A &&>(a)...); }
  template <class U> bool operator==(const RawAlloc<U> &) const noexcept { return true; }
  template <class U> bool operator!=(const RawAlloc<U> &) const noexcept { return false; }
};
typedef std::vector<u32, RawAlloc<u32>> uvec;
typedef std::vector<int32_t, RawAlloc<int32_t>> ivec; // the handle's host copies of row_ptr / col_idx (up to 6.9 GB at config 5)
// one write per page of [p, p + bytes) from all host cores (the pages of a fresh allocation are faulted in there,
// not one by one under a single-threaded copy)
void first_touch(void *p, size_t bytes);

// ---- bandk.cpp
int bandk_build(int64_t n, int64_t nnz, const u32 *rp, const u32 *adj, int k, const int *sizes,
                uvec &perm, uvec &sr_ptr, uvec &ssr_ptr);
// pieces of bandk_build that the device builder (bandk_gpu.cu) reuses: the grouping threshold and
// walk, and everything from the level-1 coarse graph on (RCM, level 2, uncoarsening)
struct CoarseGraph {
  int64_t n = 0;
  uvec rp, adj, mult;
};
u32 bandk_threshold(int size, int64_t nnz, int64_t n);
void bandk_group_rows(const u32 *rp, int64_t n, u32 T, uvec &start);
void bandk_group_rows_literal(const u32 *rp, int64_t n, u32 T, uvec &start, uvec &grp);
void bandk_group_rows_fast(const u32 *rp, int64_t n, u32 T, uvec &start, uvec &grp);
int bandk_finish(int64_t n, int k, const int *sizes, CoarseGraph &level1, uvec &map1,
                 uvec &perm, uvec &sr_ptr, uvec &ssr_ptr);
void reorder_f32(int64_t n, const u32 *perm, const u32 *rp, const u32 *ci, const float *v,
                 u32 *rp2, u32 *ci2, float *v2);
void reorder_f64(int64_t n, const u32 *perm, const u32 *rp, const u32 *ci, const double *v,
                 u32 *rp2, u32 *ci2, double *v2);

// ---- io.cpp
int read_csr_text(const char *path, int base, int dtype, int64_t *m, int64_t *n, int64_t *nnz,
                  int32_t **rp, int32_t **ci, void **val, std::string &err);
int read_csr_cached(const char *path, int base, int dtype, int64_t *m, int64_t *n, int64_t *nnz,
                    int32_t **rp, int32_t **ci, void **val, int *cache_hit, std::string &err);
int write_csr_text(const char *path, int base, int dtype, int64_t m, int64_t n, int64_t nnz,
                   const int32_t *rp, const int32_t *ci, const void *val, std::string &err);
int read_csr3_text(const char *path, int dtype, int64_t *n_ssr, int64_t *n_sr, int64_t *m,
                   int64_t *n, int64_t *nnz, u32 **ssr, u32 **sr, u32 **rp, u32 **ci, void **val,
                   std::string &err);
int write_csr3_text(const char *path, int dtype, int64_t n_ssr, int64_t n_sr, int64_t m,
                    int64_t n, int64_t nnz, const u32 *ssr, const u32 *sr, const u32 *rp,
                    const u32 *ci, const void *val, std::string &err);

// ---- partition.cpp
int partition_rows(int64_t m, const int32_t *rp, int n_parts, const u32 *align_ptr,
                   int64_t n_align, int64_t *row_lo);
int block_col_range(const int32_t *rp, const int32_t *ci, int64_t row_lo, int64_t row_hi,
                    int64_t *cmin, int64_t *cmax);
// Row-length bins of the CSR_VECTOR path.  bin_of_len() is THE rule (oracle/partition_oracle.py
// restates it): 0 empty, 1: len<=4, 2: <=8, 3: <=16, 4: <=32, 5: <=64, 6: <=BLOCK_ROW_THRESHOLD,
// 7: longer (one CTA per row).
enum { NUM_BINS = 8, BLOCK_ROW_THRESHOLD = 4096 };
inline int bin_of_len(int64_t len) {
  if (len <= 0) return 0;
  if (len <= 4) return 1;
  if (len <= 8) return 2;
  if (len <= 16) return 3;
  if (len <= 32) return 4;
  if (len <= 64) return 5;
  if (len <= BLOCK_ROW_THRESHOLD) return 6;
  return 7;
}
void row_stats(int64_t m, const int32_t *rp, int64_t out[10]);
int64_t row_len_percentile(int64_t m, const int32_t *rp, int pct);
// rows of each bin in ascending order, concatenated; bin_ptr[NUM_BINS+1]
void bin_rows(int64_t m, const int32_t *rp, std::vector<int32_t> &rows, int64_t bin_ptr[NUM_BINS + 1]);
// merge-path tile table: tile_row[t] = last row r with rp[r] <= t*tile (t = 0..n_tiles), i.e. the
// row that owns nonzero t*tile; tile_row[n_tiles] = m
void tile_rows(int64_t m, const int32_t *rp, int64_t tile, std::vector<int32_t> &tile_row);

// CSR_BLOCK row blocks: consecutive rows are gathered into a block while the block's nonzeros stay <= cap and its
// rows <= max_rows; a row that alone exceeds cap is a block of its own.  blk_row[b] = first row of block b,
// blk_row[n_blocks] = m (oracle/partition_oracle.py: block_rows restates the rule).
void block_rows(int64_t m, const int32_t *rp, int64_t cap, int64_t max_rows, std::vector<int32_t> &blk_row);

// column order of hspmv_relabel_hub_columns: descending reference count, ascending id inside a count
void hub_order(int64_t n, const u32 *cnt, uvec &perm, uvec &inv);
// plan of the pipelined host call: row chunks and the x piece every chunk waits for (partition.cpp)
void pipe_chunks(int64_t m, int64_t vs, std::vector<int64_t> &row_lo);
void pipe_pieces(int64_t n, const std::vector<int64_t> &chunk_cmax, std::vector<int64_t> &x_lo, std::vector<int> &need);

// ---- sts.cpp: level sets of the triangular solve (kernel "STS", order "LS", k = 2), bit-exact vs oracle/sts_oracle.cpp
int sts_levels(int64_t n, const int32_t *rp, const int32_t *ci, uvec &perm, std::vector<int32_t> &pack_ptr);

// ---- heuristics.cpp
int auto_params(int64_t nnz, int64_t n_rows, int arch, int sizes[2], int *veclevel, int *block_dim_y);

} // namespace hspmv
