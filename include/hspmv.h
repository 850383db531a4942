/* hspmv.h -- C ABI of the B200-native SpMV engine (libhspmv.so).
 *
 * This is the drop-in boundary for the y = A*x hot path of computablee/heterogeneous-spmv.
 * The reference has no FFI layer; its "operator API" is the set of C/C++ entry points its
 * drivers call.  Every function below names the reference interface
 * it replaces (file:line relative to the reference tree).  Plain pointers and sizes only;
 * no C++ or torch types; int status returns (0 = HSPMV_OK); opaque handles; host buffers are
 * caller-owned, device buffers are library-owned unless a function says otherwise.
 *
 * Threading: one handle per host thread.  All launches go to the stream passed in (or the
 * handle's own stream when NULL is passed where noted).  There is NO CPU fallback: every
 * compute entry point returns HSPMV_ECUDA when no sm_100 device is usable.
 *
 * include/spmv.h and include/csrk.h are source-compatible mirrors of the reference's
 * spmv-csr/spmv.h and spmv-csrk/csrk.h built on top of this ABI.
 */
#ifndef HSPMV_H
#define HSPMV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hspmv_matrix hspmv_matrix; /* opaque */

enum { HSPMV_F32 = 0, HSPMV_F64 = 1 };

enum {
  HSPMV_OK = 0,
  HSPMV_EINVAL = 1,       /* bad argument */
  HSPMV_ECUDA = 2,        /* CUDA runtime error / no usable device */
  HSPMV_EIO = 3,          /* file could not be opened / parsed */
  HSPMV_ENOMEM = 4,
  HSPMV_EUNSUPPORTED = 5, /* e.g. Band-k on a matrix with a coarse vertex without upper entries (Q8) */
  HSPMV_ESTATE = 6        /* call order: e.g. CSR-k kernel before hspmv_build_csrk */
};

/* Kernel selection.  AUTO picks from the row-length statistics gathered at creation. */
enum {
  HSPMV_ALGO_AUTO = 0,
  HSPMV_ALGO_CSR_VECTOR = 1, /* row-length binned: sub-warp vector-per-row (2..32 lanes), block-per-row */
  HSPMV_ALGO_CSR_STREAM = 2, /* fixed-row tiles staged by TMA bulk copies (CSR-k engine, virtual super-rows) */
  HSPMV_ALGO_CSR_MERGE = 3,  /* nnz-balanced split with carry fix-up (power-law matrices) */
  HSPMV_ALGO_CSRK = 4,       /* SSR -> CTA, SR -> warp, row -> lane(s), TMA-staged super-rows (needs build_csrk) */
  HSPMV_ALGO_CSR_BLOCK = 5   /* row blocks of <= HSPMV_OPT_BLOCK_NNZ nonzeros, one CTA each, products staged in shared memory:
                                matrices that live in L2 (tens of microseconds per launch) and ragged rows */
};

/* Tunables for hspmv_set_option (all optional; defaults come from the auto-tuner). */
enum {
  HSPMV_OPT_VEC = 1,          /* lanes per row: 1,2,4,8,16,32; 0 = auto */
  HSPMV_OPT_WARPS = 2,        /* warps per CTA for the TMA kernels */
  HSPMV_OPT_STAGES = 3,       /* TMA ring depth per warp */
  HSPMV_OPT_STAGE_NNZ = 4,    /* capacity of one stage in nonzeros */
  HSPMV_OPT_CTAS_PER_SM = 5,  /* persistent grid = SMs * this */
  HSPMV_OPT_ROWS_PER_TILE = 6,/* CSR_STREAM: rows per warp tile */
  HSPMV_OPT_L2_PERSIST_X = 7, /* 1: put x under a persisting L2 access-policy window */
  HSPMV_OPT_MERGE_TILE = 8,   /* CSR_MERGE: nonzeros per CTA tile */
  HSPMV_OPT_THREADS = 9,      /* CSR_VECTOR / CSR_MERGE block size */
  HSPMV_OPT_NO_PIPELINE = 11, /* 1: hspmv_spmv_host copies, computes, copies back strictly in sequence */
  HSPMV_OPT_TAIL = 13,        /* TMA kernels: masked gathers per lane per round (compiled: see csrc/csrk_variants.h); 0 = default */
  HSPMV_OPT_BANDK_HOST = 12,  /* 1: hspmv_build_csrk runs entirely on the host (default: coarse graph + reorderA in HBM) */
  HSPMV_OPT_STS_BARRIERS = 15,/* hspmv_sts_solve: 1 = level-scheduled kernel (a block / grid barrier per pack), 2 = barrier-free (per-row
                                 ready flags), 0 = by the mean pack size (barrier-free below 256 rows per pack) */
  HSPMV_OPT_STS_GROUPS = 16,  /* barrier-free STS solve: resident lane groups (0 = as many as the device takes) */
  HSPMV_OPT_BLOCK_NNZ = 14,   /* CSR_BLOCK: nonzeros per row block (256..16384; 0 = 2048) */
  HSPMV_OPT_CHUNK = 10        /* TMA kernels: unmasked x gathers a lane keeps in flight per chunk; 0 = auto, < 0 = masked loop only */
};

/* ---- runtime ---------------------------------------------------------------------- */
int hspmv_version(void);
/* last error text of the calling thread ("" if none) */
const char *hspmv_last_error(void);
int hspmv_device_count(int *count);
int hspmv_set_device(int device);
/* sm count, L2 bytes, total HBM bytes, cc major*10+minor of the current device */
int hspmv_device_info(int *sm_count, int64_t *l2_bytes, int64_t *hbm_bytes, int *cc);
/* number of kernels this library has launched in the calling process (bench.py gpu_launches) */
int64_t hspmv_launch_count(void);

/* ---- file formats ----------------------------------------------------------------- */
/* Replaces my_read_csr: spmv-csr/spmv.c:11-57 (index_base 0), spmv-csrk/spmv.cpp:32-79 and
 * cuda-spmv-csrk/cuda/spmv.cu:46-93 (index_base 1: 1 is subtracted from every index),
 * cuda-spmv-csr/spmv.cu:11-73.  Grammar: "m n nnz", (m+1) row pointers, nnz column indices,
 * nnz values, whitespace separated.  dtype selects float or double storage of *val.
 * Arrays are allocated by the library (free with hspmv_free_host).  The file is mmap'ed and parsed
 * by all host threads (OMP_NUM_THREADS); a malformed or missing token fails the call
 * (HSPMV_EIO) instead of leaving the arrays half-filled as fscanf does.  Tokens beyond the
 * announced counts are ignored, as in the reference readers. */
int hspmv_read_csr_text(const char *path, int index_base, int dtype, int64_t *m, int64_t *n,
                        int64_t *nnz, int32_t **row_ptr, int32_t **col_idx, void **val);
/* Same result as hspmv_read_csr_text, through a binary side file `<path>.hspmvbin` (64-byte header
 * with the text file's size and mtime, then the three arrays): loaded with one read per array when
 * it matches the text file, index base and dtype; otherwise the text is parsed and the side file
 * (re)written, best effort.  *cache_hit (optional) tells which happened.  The reference harness
 * re-parses the same text file on every run (run_scripts/run_cuda.py:84-92). */
int hspmv_read_csr_cached(const char *path, int index_base, int dtype, int64_t *m, int64_t *n,
                          int64_t *nnz, int32_t **row_ptr, int32_t **col_idx, void **val,
                          int *cache_hit);
/* Writes the layout of helpers/converter.m:42-48 (also reformat-csr-to-csr3/spmv.cpp:38-62). */
int hspmv_write_csr_text(const char *path, int index_base, int dtype, int64_t m, int64_t n,
                         int64_t nnz, const int32_t *row_ptr, const int32_t *col_idx,
                         const void *val);
/* `.csr3` reader; replaces reformat-csr-to-csr3/stats.c:10-79.  Header "nSSR nSR M N NNZ",
 * then ssr_ptr, sr_ptr, row_ptr, col (all 0-based, %u), val (%.6f). */
int hspmv_read_csr3_text(const char *path, int dtype, int64_t *n_ssr, int64_t *n_sr, int64_t *m,
                         int64_t *n, int64_t *nnz, uint32_t **ssr_ptr, uint32_t **sr_ptr,
                         uint32_t **row_ptr, uint32_t **col_idx, void **val);
/* `.csr3` writer; replaces my_write_csr, reformat-csr-to-csr3/spmv-auto.cpp:30-65. */
int hspmv_write_csr3_text(const char *path, int dtype, int64_t n_ssr, int64_t n_sr, int64_t m,
                          int64_t n, int64_t nnz, const uint32_t *ssr_ptr, const uint32_t *sr_ptr,
                          const uint32_t *row_ptr, const uint32_t *col_idx, const void *val);
void hspmv_free_host(void *p);

/* ---- matrix handles --------------------------------------------------------------- */
/* Upload once into device-resident, 256 B-aligned, tail-padded arrays.  Replaces the upload in
 * cuda-spmv-csr/spmv.cu:28-30,66-71 and the CSRk_Graph ctor cuda-spmv-csrk/cuda/csrk.cu:532-588
 * (which, like this, copies the caller's arrays: spmv-csrk/csrk.cpp:392-408).
 * keep_host != 0 keeps a host copy inside the handle (needed by hspmv_build_csrk). */
int hspmv_create_csr(hspmv_matrix **out, int64_t m, int64_t n, int64_t nnz,
                     const int32_t *row_ptr, const int32_t *col_idx, const void *val, int dtype,
                     int index_base, int keep_host);
/* From `.csr3` arrays (matrix already Band-k reordered); k = 3, or k = 2 when ssr_ptr == NULL. */
int hspmv_create_csrk(hspmv_matrix **out, int64_t n_ssr, int64_t n_sr, int64_t m, int64_t n,
                      int64_t nnz, const uint32_t *ssr_ptr, const uint32_t *sr_ptr,
                      const uint32_t *row_ptr, const uint32_t *col_idx, const void *val,
                      int dtype);
/* Synthetic matrices generated directly in HBM (the shapes of BASELINE.json configs; same
 * closed forms as generators.py): rows [row_lo,row_hi) of the g^3 27-point stencil (kind 27)
 * or the g^2 5-point Laplacian (kind 5), natural order; row_ptr rebased to 0; column ids
 * shifted by -col_shift (multi-GPU localisation).  mode 0 = classic values, 1 = varied. */
int hspmv_create_stencil(hspmv_matrix **out, int kind, int64_t g, int dtype, int mode,
                         int64_t row_lo, int64_t row_hi, int64_t col_shift, int64_t n_cols);
int hspmv_destroy(hspmv_matrix *A);

/* m, n, nnz, dtype, k (1 = plain CSR), #super-rows, #super-super-rows; any pointer may be NULL */
int hspmv_matrix_info(const hspmv_matrix *A, int64_t *m, int64_t *n, int64_t *nnz, int *dtype,
                      int *k, int64_t *n_sr, int64_t *n_ssr);
/* Device -> host copy of the (possibly Band-k reordered) CSR arrays, for bit-exact diffing. */
int hspmv_download_csr(const hspmv_matrix *A, int32_t *row_ptr, int32_t *col_idx, void *val);
/* Row-length statistics and the bins the CSR_VECTOR path uses: out[0..7] = rows per bin
 * (empty, <=4, <=8, <=16, <=32, <=64, <=128 [vec 32 loops], >block_row_threshold),
 * out[8] = max row length, out[9] = number of empty rows. */
int hspmv_row_stats(const hspmv_matrix *A, int64_t out[10]);

/* ---- Band-k / CSR-k --------------------------------------------------------------- */
/* Replaces CSRk_Graph::putInCSRkFormat (spmv-csrk/csrk.cpp:681-706 = BAND_k::
 * preprocessingForSpMV :841-1067 + reorderA :548-676) and the device upload of the maps
 * (cuda-spmv-csrk/cuda/csrk.cu:848-866).  k in {2,3}; sup_row_sizes[0] controls level 1 (rows
 * per super-row, nnz-weighted), [1] level 2 -- same order as the reference argv (Q9).
 * Unlike the reference CUDA path (quirk Q3) the REORDERED matrix is what gets uploaded. */
int hspmv_build_csrk(hspmv_matrix *A, int k, const int *sup_row_sizes);
/* Host copies for bit-exact diffing; any pointer may be NULL.  perm[m] (new row -> old row,
 * CSRk_Graph::getPermutation csrk.h:325), sr_ptr[n_sr+1], ssr_ptr[n_ssr+1]. */
int hspmv_csrk_get(const hspmv_matrix *A, uint32_t *perm, uint32_t *sr_ptr, uint32_t *ssr_ptr);

/* The same index work without a device (host only; used by tests and by converters). */
int hspmv_bandk_host(int64_t n, int64_t nnz, const uint32_t *row_ptr, const uint32_t *col_idx,
                     int k, const int *sup_row_sizes, uint32_t *perm, uint32_t *sr_ptr,
                     int64_t *n_sr, uint32_t *ssr_ptr, int64_t *n_ssr);
int hspmv_reorder_host(int64_t n, int64_t nnz, const uint32_t *perm, const uint32_t *row_ptr,
                       const uint32_t *col_idx, const void *val, int dtype, uint32_t *row_ptr_out,
                       uint32_t *col_idx_out, void *val_out);

/* Size / launch-shape heuristics.  arch: 0 = Volta formula (spmv-auto-volta.cu:130-159 and
 * reformat-csr-to-csr3/spmv-auto.cpp:154-173), 1 = Ampere (spmv-auto-ampere.cu:130-158),
 * 2 = MI100 (hip/spmv-auto-mi100.cu:132-133), 3 = B200 retune (this library).
 * Outputs sizes[0] ("ssrs"), sizes[1] ("srs"), veclevel (0 = scalar kernel), block_dim_y. */
int hspmv_auto_params(int64_t nnz, int64_t n_rows, int arch, int sizes[2], int *veclevel,
                      int *block_dim_y);

/* ---- vectors ---------------------------------------------------------------------- */
/* Replaces CSRk_Graph::setX (csrk.h:327-335, csrk.cu:91-101): x_test[i] = x[perm[i]] when the
 * matrix is in CSR-k order, then H2D.  x has n entries of the matrix dtype. */
int hspmv_set_x(hspmv_matrix *A, const void *x_host);
/* Replaces getY (csrk.cu:109-112) / the D2H at cuda-spmv-csr/spmv.cu:270.  original_order != 0
 * undoes the Band-k row permutation (y_out[perm[i]] = y[i]); 0 returns the kernel's order,
 * which is what spmv-csrk/spmv.cpp:203 compares against y_csrser[perm[i]]. */
int hspmv_get_y(hspmv_matrix *A, void *y_host, int original_order);
/* Replaces CSRk_Graph::setY (cuda-spmv-csrk/cuda/csrk.cu:103-107): uploads m entries into the
 * handle's device y as they are (kernel row order). */
int hspmv_set_y(hspmv_matrix *A, const void *y_host);
/* Library-owned device vectors of the handle (x in kernel order, n entries; y, m entries). */
void *hspmv_x_device(hspmv_matrix *A);
void *hspmv_y_device(hspmv_matrix *A);
/* Wall seconds of the last hspmv_build_csrk that ran the device builder, by phase (up to n = 6 entries): 0 host
 * preparation (row_ptr download, row statistics, greedy grouping -- csrk.cpp:1278-1292), 1 level-1 coarse graph in HBM
 * and its download (handCoarsen, :1243-1437), 2 RCM / level 2 / uncoarsening on the host cores (:2289-2475, :1148-1224),
 * 3 A' = P A P^T in HBM (reorderA, :548-676), 4 CSR-k maps and statistics of the new matrix, 5 total.  All zero when
 * the host builder ran.  bench.py prints them beside `bandk_build_s`. */
int hspmv_build_phases(const hspmv_matrix *A, double *seconds, int n);
/* The device arrays behind a handle, for callers that launch their OWN kernels on them the way the reference's
 * CUDA drivers do -- cuda-spmv-csrk/cuda/spmv.cu:33-44 declares these as `extern` device-pointer globals
 * (r_vec_gpu, c_vec_gpu, val_gpu, x_test_gpu, y_gpu, mapCoarseToFinerRows_gpu_outer / _inner, and the device WORDS
 * N_gpu, M_gpu, NNZ_gpu, numCoarsestRows_gpu) that csrk.cu:putInCSRkFormat fills; include/csrk.cuh binds them from
 * this view.  All pointers are library-owned device memory, valid until the handle is destroyed or reordered
 * (hspmv_build_csrk / hspmv_build_sts / hspmv_relabel_hub_columns replace the matrix arrays: fetch the view again).
 * sr_ptr / ssr_ptr are NULL when the level does not exist; dims = device uint32[4] {n_rows, n_cols, nnz, n_coarsest}
 * with n_coarsest = n_ssr (k = 3), n_sr (k = 2) or n_rows (k = 1). */
typedef struct hspmv_device_view {
  void *row_ptr, *col_idx, *val, *x, *y; /* uint32[m+1], uint32[nnz], T[nnz], T[n], T[m] */
  void *sr_ptr, *ssr_ptr;                /* uint32[n_sr+1], uint32[n_ssr+1] */
  void *dims;                            /* device uint32[4] */
  int64_t n_coarsest;
  int dtype, k;
} hspmv_device_view;
int hspmv_device_view_get(hspmv_matrix *A, hspmv_device_view *out);
/* x_i = 1 + ((global_offset + i) mod 97)/128 written straight into the handle's device x
 * (generators.py:vector_x; the non-constant x the parity tests use instead of the reference's
 * x = 1, which is blind to permutation errors -- SURVEY.md section 4). */
int hspmv_fill_x_synthetic(hspmv_matrix *A, int64_t global_offset);
/* The same rule times `scale`, on up to 8 index ranges [lo[s], hi[s]) of any device vector, in one
 * stream-ordered launch: x_i = (1 + ((global_offset + i) mod 97)/128) * scale.  The multi-GPU parity
 * runs use it to change the boundary slices of x every step, so that every step needs a fresh halo
 * (a push that never lands, or lands at the wrong offset, then shows in y). */
int hspmv_fill_x_segments(void *d_x, int dtype, int n_seg, const int64_t *lo, const int64_t *hi,
                          int64_t global_offset, double scale, void *stream);

/* ---- the hot path ----------------------------------------------------------------- */
int hspmv_set_option(hspmv_matrix *A, int option, int64_t value);
/* Enqueue y = A*x on `stream` (a cudaStream_t cast to void*; NULL = the legacy default
 * stream).  d_x / d_y are device pointers in kernel order; NULL selects the handle's own
 * vectors.  Replaces the launches at cuda-spmv-csr/spmv.cu:243, cuda-spmv-csrk/cuda/spmv.cu:
 * 186-198, spmv-auto-ampere.cu:205-217 and the CPU loops omp_spmv (spmv-csr/spmv.c:92-114)
 * and CSRk_Graph::SpMV (spmv-csrk/csrk.cpp:149-330). */
int hspmv_spmv(hspmv_matrix *A, int algo, const void *d_x, void *d_y, void *stream);
/* Same, restricted to rows [row_lo,row_hi) (CSR_STREAM / CSR_VECTOR on a plain-CSR handle): lets
 * the multi-GPU layer run the interior rows while the x halo is still in flight. */
int hspmv_spmv_range(hspmv_matrix *A, int algo, int64_t row_lo, int64_t row_hi, const void *d_x,
                     void *d_y, void *stream);
/* Blocks until everything this handle has enqueued (on its own stream and on the stream of the
 * last hspmv_spmv) has finished: the cudaDeviceSynchronize() after each launch in the reference
 * drivers (cuda-spmv-csr/spmv.cu:248, cuda-spmv-csrk/cuda/spmv.cu:200). */
int hspmv_sync(hspmv_matrix *A);
/* N timed launches after `warmup` untimed ones; ms_out[i] = device time of launch i from
 * CUDA events on the launching stream.  flush_l2 = 1 overwrites a scratch buffer of twice the L2
 * size before every launch (outside the event pair); L2 is then full of DIRTY foreign lines, whose
 * write-back the timed launch also pays for.  flush_l2 = 2 additionally reads a second such buffer
 * after the write, so that L2 is full of CLEAN foreign lines: the cold-cache time of the product
 * alone.  Replaces the host-timer loops at
 * cuda-spmv-csr/spmv.cu:240-260 / spmv-auto-ampere.cu:205-230. */
int hspmv_spmv_timed(hspmv_matrix *A, int algo, int warmup, int iters, int flush_l2,
                     float *ms_out);
/* `iters` launches back to back between ONE pair of events after `warmup` untimed ones (the
 * measurement SURVEY.md 8(d) asks for; per-launch event pairs add ~5 us of launch latency to each
 * sample, which matters for the 30-60 us launches of configs 1-2).  *ms_per_launch = total / iters. */
int hspmv_spmv_timed_batch(hspmv_matrix *A, int algo, int warmup, int iters, float *ms_per_launch);
/* B200 autotuner (the reference tunes by sweeping its drivers from
 * run_scripts/run_tuning.py:89 and by the per-architecture formulas of spmv-auto-*.cu:130-158).
 * Times the compiled kernel shapes around the heuristic choice for THIS matrix (lanes per row,
 * unmasked chunk vs masked one-round tail, rows per tile) with hspmv_spmv_timed_batch and leaves
 * the fastest in the handle's options.  effort 1: ~10 candidates, 2: ~25.  The x / y contents
 * are used as they are; y is overwritten.  *best_ms = ms per launch of the winner. */
int hspmv_autotune(hspmv_matrix *A, int algo, int effort, float *best_ms);
/* End-to-end plugin call with HOST buffers: set_x (permute + H2D) -> kernel -> get_y (D2H,
 * original row order), synchronous.  This is setX + SpMV + getY of the reference drivers.
 * On large plain-CSR handles the three stages are pipelined over row chunks on three streams:
 * x pieces go up in order, a row chunk is launched (hspmv_spmv_range) as soon as the last x piece
 * it references has landed, its y piece goes down while later chunks compute.  Needs pinned host
 * buffers to overlap; pageable buffers still work, without the overlap. */
int hspmv_spmv_host(hspmv_matrix *A, int algo, const void *x_host, void *y_host);
/* `iters` back-to-back hspmv_spmv_host calls bracketed by CUDA events on the handle's stream:
 * *ms_total covers every H2D copy of x, permutation, kernel and D2H copy of y. */
int hspmv_spmv_host_timed(hspmv_matrix *A, int algo, const void *x_host, void *y_host, int iters,
                          float *ms_total);
/* Solver-style loops on small matrices (a product of tens of microseconds): `iters` products on the handle's own
 * x / y captured ONCE as a CUDA graph and replayed with one host call per `iters` products.  Replaces the reference
 * drivers' launch + cudaDeviceSynchronize per product (cuda-spmv-csr/spmv.cu:240-260).  Options are frozen at capture;
 * hspmv_set_option drops the graph. */
int hspmv_spmv_graph(hspmv_matrix *A, int algo, int iters, void *stream);
int hspmv_spmv_graph_timed(hspmv_matrix *A, int algo, int iters, int reps, float *ms_per_product);
/* Name of the kernel variant the last hspmv_spmv on this handle launched (for reports). */
const char *hspmv_last_kernel(const hspmv_matrix *A);

/* Source-compatible replacements for spmv-csr/spmv.h:11-12 (double) and the float body in
 * spmv-csr/spmv.c:92-114: host CSR in, host y out.  The matrix is uploaded on first use and
 * cached by (row_start, col_idx, val, m, nnz) identity plus a content fingerprint (all of
 * row_start, 4096 samples of col_idx / val) that is re-checked on every call, so repeated calls
 * in a timing loop only move x and y, while arrays that were refilled in place or reallocated at
 * the same addresses are uploaded again. */
int hspmv_omp_spmv_f64(int m, int n, int nnz, const int *row_start, const int *col_idx,
                       const double *val, const double *x, double *y);
int hspmv_omp_spmv_f32(int m, int n, int nnz, const int *row_start, const int *col_idx,
                       const float *val, const float *x, float *y);
void hspmv_omp_spmv_cache_clear(void);
int hspmv_omp_spmv_cache_size(void); /* matrices currently resident in that cache (<= 4, LRU) */

/* ---- multi-GPU index work (host; bit-exact vs oracle/partition_oracle.py) ---------- */
/* nnz-balanced contiguous row blocks: row_lo[p] = first row r with row_ptr[r] >= p*nnz/P
 * (row_lo[0] = 0, row_lo[P] = m).  If align_ptr != NULL (sr_ptr or the composed ssr->row map,
 * n_align+1 entries) each cut is moved up to the next group boundary. */
int hspmv_partition_rows(int64_t m, const int32_t *row_ptr, int n_parts, const uint32_t *align_ptr,
                         int64_t n_align, int64_t *row_lo);
/* For the row block [row_lo,row_hi): smallest and largest column index referenced. */
int hspmv_block_col_range(const int32_t *row_ptr, const int32_t *col_idx, int64_t row_lo,
                          int64_t row_hi, int64_t *col_min, int64_t *col_max);
/* Device variant for matrices that only exist in HBM (generated or already uploaded). */
int hspmv_col_range_device(const hspmv_matrix *A, int64_t *col_min, int64_t *col_max);


/* New handle with rows [row_lo,row_hi) of A, copied inside HBM; columns [col_lo, col_lo + n_cols)
 * become [0, n_cols) (n_cols <= 0: keep all columns, col_lo = 0).  HSPMV_EINVAL if the block
 * references a column outside the window.  For a CSR-k handle the bounds must be super-row
 * boundaries (k = 2) / super-super-row boundaries (k = 3) -- what hspmv_partition_rows returns with
 * align_ptr -- and the slice is again CSR-k, with sr_ptr / ssr_ptr rebased; x and y of a slice are
 * in A's (reordered) numbering, there is no permutation attached.  This is how one rank's shard of
 * a Band-k-reordered matrix is cut out. */
int hspmv_slice_rows(const hspmv_matrix *A, int64_t row_lo, int64_t row_hi, int64_t col_lo, int64_t n_cols,
                     hspmv_matrix **out);
/* Contiguous range [row_a,row_b) of rows whose column indices all lie in [own_col_lo,own_col_hi)
 * (local column ids): row_a = 1 + last row of the first half that references a column outside,
 * row_b = first such row of the second half.  Rows outside the range wait for the x halo. */
int hspmv_interior_rows(const hspmv_matrix *A, int64_t own_col_lo, int64_t own_col_hi,
                        int64_t *row_a, int64_t *row_b);

/* ---- NVLink halo exchange over peer memory (one process per GPU) --------------------------
 * The reference has no multi-GPU code; this replaces what would otherwise be NCCL send/recv.
 * A buffer that came from cudaMalloc (hspmv_x_device(), hspmv_peer_alloc()) is exported as a
 * 64-byte CUDA-IPC handle, opened by the neighbour, and written directly with copy-engine
 * copies; a flag word per sender publishes the step.  Everything is stream-ordered; there is no
 * host synchronisation inside a step (protocol: heterogeneous-spmv_b200/dist.py, PeerHalo). */
int hspmv_peer_export(void *dptr, void *handle64);
int hspmv_peer_open(const void *handle64, void **dptr);
int hspmv_peer_close(void *dptr);
int hspmv_peer_alloc(int64_t bytes, void **dptr);
int hspmv_peer_free(void *dptr);
/* Front-end (copy-engine / command-processor) versions that use no SM, so they run next to a
 * persistent kernel that occupies every SM: a peer copy, a 32-bit flag write and a >= wait, all
 * stream-ordered (cuStreamWriteValue32 / cuStreamWaitValue32 resolved at run time). */
int hspmv_memcpy_async(void *dst, const void *src, int64_t bytes, void *stream);
int hspmv_stream_write(void *stream, int *addr, int value);
int hspmv_stream_wait_ge(void *stream, const int *addr, int value);
/* One launch for a whole shard: tiles of the interior rows [interior_lo, interior_hi) first, then
 * the rows below and above, which reference the x halo; a warp reaching its first halo tile spins
 * (bounded) until every wait_flags[i] >= wait_step, and the halo tiles gather x through L2
 * (ld.global.cg), not through the read-only path, because a peer writes that part of x while the
 * kernel runs.  `timed_out` points to a block of HSPMV_HALO_DIAG_WORDS (10) device ints:
 * [0] set to 1 when a wait gave up, [1] longest warp spin (1024-cycle units), [2] number of
 * warp-level waits, [3] their total spin, [4..9] a histogram of the spins.  With n_wait == 0 it
 * is hspmv_spmv over the three ranges in that order and timed_out may be NULL. */
#define HSPMV_HALO_DIAG_WORDS 10
int hspmv_spmv_halo(hspmv_matrix *A, int64_t interior_lo, int64_t interior_hi, const void *d_x, void *d_y,
                    const int *wait_flags, int n_wait, int wait_step, int *timed_out, void *stream);

/* ---- multi-GPU in ONE process (csrc/mg.cu) -----------------------------------------------------
 * "multi-GPU init (device list) + partition query ... one stream + comm per device".  For C/C++
 * drivers in the style of cuda-spmv-csrk/cuda/spmv.cu:111-266 (one host thread, default device) that want every
 * GPU of the box: the matrix is built on devices[0] (uploaded or generated, optionally Band-k-reordered there),
 * cut into nnz-balanced row blocks (CSR-k: at super-row / super-super-row boundaries) and every device keeps its
 * block with the x window it references.  Banded matrices exchange the x halo over NVLink every step (peer
 * copies + flag words, the local kernel waits for the flags itself: hspmv_spmv_halo); matrices whose blocks
 * reference all of x, or whose exchange plan is not symmetric, run with x replicated.  x and y cross this
 * boundary in the ORIGINAL numbering (the Band-k permutation is applied inside, like setX / the check at
 * spmv-csrk/spmv.cpp:203).  One handle per host thread; every call returns after ENQUEUEING unless it says so. */
typedef struct hspmv_mg hspmv_mg;
const char *hspmv_mg_last_error(void);
/* k = 1: plain CSR; k = 2 / 3: hspmv_build_csrk(k, sup_row_sizes) on devices[0] before the cut. */
int hspmv_mg_create_csr(hspmv_mg **out, const int *devices, int n_dev, int64_t m, int64_t n, int64_t nnz,
                        const int32_t *row_ptr, const int32_t *col_idx, const void *val, int dtype, int index_base,
                        int k, const int *sup_row_sizes);
/* the closed-form stencils of hspmv_create_stencil (kind 27 / 5), generated in HBM on devices[0] */
int hspmv_mg_create_stencil(hspmv_mg **out, const int *devices, int n_dev, int kind, int64_t g, int dtype, int mode,
                            int k, const int *sup_row_sizes);
int hspmv_mg_destroy(hspmv_mg *M);
int hspmv_mg_info(const hspmv_mg *M, int *n_dev, int64_t *m, int64_t *n, int64_t *nnz, int *dtype, int *k,
                  int *replicated);
/* Partition query (any pointer may be NULL): devices[n_dev], row_lo[n_dev+1] (kernel row numbering), the column
 * window [col_min[d], col_max[d]] of each block, its nonzeros, and the x entries it receives per step. */
int hspmv_mg_partition_query(const hspmv_mg *M, int *devices, int64_t *row_lo, int64_t *col_min, int64_t *col_max,
                             int64_t *nnz_per_dev, int64_t *halo_elems);
int hspmv_mg_get_perm(const hspmv_mg *M, uint32_t *perm); /* CSR-k: new -> old, m entries */
void *hspmv_mg_shard(hspmv_mg *M, int d);                 /* the hspmv_matrix* of block d (options, last_kernel) */
/* Upload x (n entries, original numbering): every device receives the slice it OWNS; halo entries only ever
 * arrive device to device inside hspmv_mg_spmv.  Replicated mode: all of x to every device.  Synchronous. */
int hspmv_mg_set_x(hspmv_mg *M, const void *x_host);
/* One step on every device (halo exchange of the current x + local SpMV), enqueued only. */
int hspmv_mg_spmv(hspmv_mg *M);
/* Wait for everything enqueued; HSPMV_ECUDA if an in-kernel halo wait gave up. */
int hspmv_mg_sync(hspmv_mg *M);
int hspmv_mg_get_y(hspmv_mg *M, void *y_host); /* m entries, original row order; synchronous */
/* ms_per_step[i] = longest device time of step i (host-synchronised between steps, like the reference drivers'
 * loop at cuda-spmv-csr/spmv.cu:240-260); _batch: iters steps back to back, one mean. */
int hspmv_mg_spmv_timed(hspmv_mg *M, int warmup, int iters, float *ms_per_step);
int hspmv_mg_spmv_timed_batch(hspmv_mg *M, int warmup, int iters, float *ms_per_step);
/* test hook: NaN into every halo region of both x buffers (only a correct push makes them finite again) */
int hspmv_mg_poison_halos(hspmv_mg *M);

/* ---- power-law matrices: hub-column relabel ---------------------------------------------------------
 * Renumbers the COLUMNS of a plain-CSR handle by how often they are referenced (most popular first; ties by column
 * id), in HBM.  The x entries the gathers hit most become neighbours in memory, so the gathers of an R-MAT-like matrix
 * are served by L1 / L2 instead of pulling one 32-byte sector per 8-byte entry from all over x.  Rows are not moved
 * (y keeps its numbering, the column order inside a row is no longer sorted); hspmv_set_x and hspmv_spmv_host apply
 * the permutation to x on the way in, a caller that passes its own device x must supply x_kernel[i] = x[perm[i]].
 * perm_out (optional): n entries, new column -> old column.  One-time set-up, like hspmv_build_csrk. */
int hspmv_relabel_hub_columns(hspmv_matrix *A, uint32_t *perm_out);

/* ---- STS: sparse triangular solve on CSR-k, k = 2, level-set ordering --------------------
 * Replaces, for kernel "STS" / order "LS" / k = 2: BAND_k::preprocessingForSTS (spmv-csrk/csrk.cpp:1522-1672,
 * find_levels :2704-2816, levels sorted by size), CSRk_Graph::reorderA, incomplete_choloskey (:708-789), compute_b
 * (:791-809) and lowerSTS (:79-89).  COLOR ordering (Boost sequential_vertex_coloring) and k >= 3 are out of scope.
 * hspmv_build_sts: level sets on the host (bit-exact permutation and pack bounds), A' = P A P^T in HBM; L is the part
 * of A' up to and including the diagonal of every row (not materialised).  Afterwards the handle is a plain-CSR
 * handle of A' (k stays 1; set_x / get_y apply the permutation), and the STS entry points below work on it. */
int hspmv_sts_levels_host(int64_t n, int64_t nnz, const int32_t *row_ptr, const int32_t *col_idx, uint32_t *perm,
                          int32_t *pack_ptr, int64_t *n_packs); /* the index work alone; pack_ptr: room for n + 1 */
int hspmv_build_sts(hspmv_matrix *A);
int hspmv_sts_get(const hspmv_matrix *A, uint32_t *perm, int32_t *pack_ptr, int64_t *n_packs);
int hspmv_sts_compute_b(hspmv_matrix *A); /* b = L * 1 (compute_b), left to right per row */
/* original_order = 0: the solver's (reordered) numbering, as the reference's b / x arrays; 1: the input matrix' */
int hspmv_sts_set_b(hspmv_matrix *A, const void *b_host, int original_order);
int hspmv_sts_get_b(hspmv_matrix *A, void *b_host, int original_order);
/* lowerSTS in one cooperative launch, in one of two forms (HSPMV_OPT_STS_BARRIERS; default: chosen by the mean pack
 * size).  Level-scheduled: packs in order with a block / grid barrier between them.  Barrier-free: rows dealt to lane
 * groups in increasing order, every row waits (bounded spin) for the ready flags of exactly the x entries it reads, so
 * consecutive packs overlap. */
int hspmv_sts_solve(hspmv_matrix *A, void *stream);
int hspmv_sts_status(hspmv_matrix *A, int *timed_out); /* synchronises; *timed_out = 1 if a dependency wait ever gave up */
int hspmv_sts_get_x(hspmv_matrix *A, void *x_host, int original_order);
int hspmv_sts_solve_timed(hspmv_matrix *A, int warmup, int iters, float *ms_out);

/* ---- kernel-side index tables, exposed for the bit-exactness tests ------------------ */
/* CSR_MERGE tile table: tile_row[t] = last row r with row_ptr[r] <= t*tile, t = 0..n_tiles
 * (n_tiles = ceil(nnz/tile)); tile_row[n_tiles] = m.  Host rule and the device copy. */
int hspmv_merge_tile_rows_host(int64_t m, const int32_t *row_ptr, int64_t tile, int32_t *tile_row);
int hspmv_merge_tile_rows_device(const hspmv_matrix *A, int64_t *tile, int64_t *n_tiles,
                                 int32_t *tile_row);
/* Smallest L such that at least pct % of the rows have <= L entries (the 90th percentile sizes the
 * masked kernel shape for ragged matrices). */
int64_t hspmv_row_len_percentile_host(int64_t m, const int32_t *row_ptr, int pct);
/* CSR_BLOCK row blocks: greedy runs of consecutive rows with <= cap nonzeros and <= max_rows rows (a longer row is
 * alone); blk_row needs room for m + 2 entries. */
int hspmv_block_rows_host(int64_t m, const int32_t *row_ptr, int64_t cap, int64_t max_rows, int32_t *blk_row,
                          int64_t *n_blocks);
/* Band-k level grouping (handCoarsen's greedy walk, spmv-csrk/csrk.cpp:1278-1292): start[] of the groups (room for
 * n + 2 entries), *n_starts entries written, grp[n] = group of every row.  literal != 0 runs the row-by-row statement
 * of the rule, 0 the forward scan of row_ptr that hspmv_build_csrk uses; the tests compare the two. */
int hspmv_bandk_group_rows_host(int64_t n, const uint32_t *row_ptr, uint32_t threshold, int literal, uint32_t *start,
                                int64_t *n_starts, uint32_t *grp);
/* Column order of hspmv_relabel_hub_columns, host rule: columns by descending reference count, ascending id inside a
 * count; perm[new] = old, inv[old] = new (n entries each).  Exposed for the tests. */
int hspmv_hub_order_host(int64_t n, const uint32_t *count, uint32_t *perm, uint32_t *inv);
/* Plan of the pipelined host call (hspmv_spmv_host on a large plain-CSR handle): equal row chunks row_lo[0..C] (~4 MB
 * of y each, 4 <= C <= 32) and -- when chunk_cmax[C] (highest column each chunk references, -1 = none) is given -- the
 * x pieces x_lo[0..C], cut where the chunks' column ranges end so that a banded chunk waits for its own piece only,
 * and need[c] = the last piece chunk c waits for.  row_lo / x_lo need room for 66 entries, need for 65.  Host rule
 * behind the call the bench's `e2e` leg times; exposed for the property tests. */
int hspmv_pipe_plan_host(int64_t m, int64_t n, int vsize, const int64_t *chunk_cmax, int64_t *row_lo, int *n_chunks, int64_t *x_lo,
                         int *need);
/* CSR_VECTOR bins: rows[m] grouped by bin in ascending row order, bin_ptr[9]. */
int hspmv_bin_rows_host(int64_t m, const int32_t *row_ptr, int32_t *rows, int64_t *bin_ptr);

#ifdef __cplusplus
}
#endif
#endif /* HSPMV_H */
