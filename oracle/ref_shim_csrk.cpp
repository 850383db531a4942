// TEST INFRASTRUCTURE ONLY -- C entry points around the UNMODIFIED reference class
// CSRk_Graph (/root/reference/spmv-csrk/csrk.h:253-345, csrk.cpp) so that tests can run
// the real Band-k builder and the real OpenMP CSR-k SpMV and read their arrays back.
// Built by oracle/Makefile into oracle/_ref/libref_csrk_{f32,f64}.so; never linked into
// the product.  `float` below is rewritten to `double` by ref_prelude_cxx.h in the f64 build.
#define private public   // read-only peeking at r_vec / c_vec / val / mapCoarseToFinerRows
#include "csrk.h"
#undef private

extern "C" {

// ctor (copies the caller's arrays, csrk.cpp:392-408) + putInCSRkFormat (csrk.cpp:681-706)
void *ref_csrk_create(long n, long nnz, unsigned int *r, unsigned int *c, float *v, int k,
                      int *sup_row_sizes) {
  CSRk_Graph *g = new CSRk_Graph(n, n, nnz, r, c, v, "SpMV", "", "HAND", false, k,
                                 sup_row_sizes);
  g->putInCSRkFormat();
  return g;
}

long ref_csrk_num_coarsest(void *h) { return ((CSRk_Graph *)h)->numCoarsestRows; }

// number of entries (minus one) of map[level]; level = k-1 is the outermost map.
long ref_csrk_map_len(void *h, int level) {
  CSRk_Graph *g = (CSRk_Graph *)h;
  if (level == g->k - 1)
    return g->numCoarsestRows;
  // the next-finer map has map[level+1][last] groups
  return g->mapCoarseToFinerRows[level + 1][ref_csrk_map_len(h, level + 1)];
}

void ref_csrk_get_map(void *h, int level, unsigned int *out) {
  CSRk_Graph *g = (CSRk_Graph *)h;
  long len = ref_csrk_map_len(h, level) + 1;
  for (long i = 0; i < len; i++)
    out[i] = g->mapCoarseToFinerRows[level][i];
}

void ref_csrk_get_matrix(void *h, unsigned int *perm, unsigned int *r, unsigned int *c,
                         float *v) {
  CSRk_Graph *g = (CSRk_Graph *)h;
  if (perm && g->permBigG)
    for (long i = 0; i < g->N; i++) perm[i] = g->permBigG[i];
  if (r) for (long i = 0; i <= g->N; i++) r[i] = g->r_vec[i];
  if (c) for (long i = 0; i < g->NNZ; i++) c[i] = g->c_vec[i];
  if (v) for (long i = 0; i < g->NNZ; i++) v[i] = g->val[i];
}

void ref_csrk_setx(void *h, float *x) { ((CSRk_Graph *)h)->setX(x); }

// y comes back in the permuted row order, as in spmv-csrk/spmv.cpp:203
void ref_csrk_spmv(void *h, float *y) { ((CSRk_Graph *)h)->SpMV(y); }

double ref_csrk_spmv_timed(void *h, float *y, int warmup, int iters) {
  CSRk_Graph *g = (CSRk_Graph *)h;
  for (int i = 0; i < warmup; i++) g->SpMV(y);
  double best = 1e30;
  for (int i = 0; i < iters; i++) {
    double t0 = omp_get_wtime();
    g->SpMV(y);
    double t = omp_get_wtime() - t0;
    if (t < best) best = t;
  }
  return best;
}

void ref_csrk_destroy(void *h) { delete (CSRk_Graph *)h; }

// ---- STS (csrk.cpp:79-89 lowerSTS k == 2, :1522-1672 preprocessingForSTS with LS ordering, :708-809) ----
// ctor("STS", "LS", "HAND", k = 2) + putInCSRkFormat (find_levels, levels sorted by size, reorderA)
// + incomplete_choloskey (L / U structure) + compute_b (b = L * 1, x = 0).
void *ref_sts_create(long n, long nnz, unsigned int *r, unsigned int *c, float *v) {
  int sizes[1] = {0};
  CSRk_Graph *g = new CSRk_Graph(n, n, nnz, r, c, v, "STS", "LS", "HAND", false, 2, sizes);
  g->putInCSRkFormat();
  g->incomplete_choloskey();
  g->compute_b();
  return g;
}
int ref_sts_num_packs(void *h) { return ((CSRk_Graph *)h)->numPacks; }
long ref_sts_nnz_l(void *h) { return ((CSRk_Graph *)h)->num_edges_L[((CSRk_Graph *)h)->N]; }
void ref_sts_get(void *h, unsigned int *perm, int *pack_ptr, unsigned int *l_rp, unsigned int *l_ci, float *l_val, float *b) {
  CSRk_Graph *g = (CSRk_Graph *)h;
  if (perm) for (long i = 0; i < g->N; i++) perm[i] = g->permBigG[i];
  if (pack_ptr) for (int i = 0; i <= g->numPacks; i++) pack_ptr[i] = g->packsPointer[i];
  if (l_rp) for (long i = 0; i <= g->N; i++) l_rp[i] = g->num_edges_L[i];
  const long nl = g->num_edges_L[g->N];
  if (l_ci) for (long i = 0; i < nl; i++) l_ci[i] = g->adj_L[i];
  if (l_val) for (long i = 0; i < nl; i++) l_val[i] = g->val_L[i];
  if (b) for (long i = 0; i < g->N; i++) b[i] = g->b[i];
}
void ref_sts_set_b(void *h, float *b) {
  CSRk_Graph *g = (CSRk_Graph *)h;
  for (long i = 0; i < g->N; i++) g->b[i] = b[i];
}
// lowerSTS(); x comes back in the REORDERED numbering (the numbering the reference solves in)
void ref_sts_solve(void *h, float *x) {
  CSRk_Graph *g = (CSRk_Graph *)h;
  for (long i = 0; i < g->N; i++) g->x[i] = 0;
  g->lowerSTS();
  for (long i = 0; i < g->N; i++) x[i] = g->x[i];
}
double ref_sts_solve_timed(void *h, int warmup, int iters) {
  CSRk_Graph *g = (CSRk_Graph *)h;
  for (int i = 0; i < warmup; i++) g->lowerSTS();
  double best = 1e30;
  for (int i = 0; i < iters; i++) {
    double t0 = omp_get_wtime();
    g->lowerSTS();
    double t = omp_get_wtime() - t0;
    if (t < best) best = t;
  }
  return best;
}

int ref_csrk_sizeof_real(void) { return (int)sizeof(float); }
}
