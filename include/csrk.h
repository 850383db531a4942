// csrk.h -- source-compatible mirror of the reference's CSRk_Graph public interface
// (spmv-csrk/csrk.h:253-345; CUDA flavour cuda-spmv-csrk/cuda/csrk.cuh:341-348) on top of the
// C ABI in hspmv.h.  Same constructor arguments, same method names, same ownership rules: the
// constructor copies the caller's CSR arrays (spmv-csrk/csrk.cpp:392-408), getPermutation()
// returns a pointer owned by the object.
//
//   val_type is float, as in the reference (csrk.h:40); compile with -DHSPMV_VAL_DOUBLE for the
//   fp64 interface north_star names as primary.
//
// Differences that a maintainer should know about (all deliberate):
//   * putInCSRkFormat() uploads the REORDERED matrix (the reference CUDA path uploads before
//     reordering -- quirk Q3);
//   * only kernelCalled == "SpMV" and inCoarsenType == "HAND" are accepted (every reference
//     driver hard-codes them, spmv-csrk/spmv.cpp:124-125); anything else prints and exits(1) like
//     the reference's "Undefined kernel invoked" path (csrk.cpp:699-701).
#ifndef HSPMV_COMPAT_CSRK_H
#define HSPMV_COMPAT_CSRK_H
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "hspmv.h"

#ifdef HSPMV_VAL_DOUBLE
typedef double val_type;
#define HSPMV_VAL_DTYPE HSPMV_F64
#else
typedef float val_type;
#define HSPMV_VAL_DTYPE HSPMV_F32
#endif

class CSRk_Graph {
  hspmv_matrix *A_;
  long N_, M_, NNZ_;
  int k_;
  std::vector<int> sizes_;
  std::vector<unsigned int> perm_;
  val_type *y_;

  // include/csrk.cuh (the CUDA flavour) points the reference's device-pointer globals at the handle's arrays here
  void bind_() {
#ifdef HSPMV_CSRK_BIND_DEVICE_GLOBALS
    HSPMV_CSRK_BIND_DEVICE_GLOBALS(A_);
#endif
  }

  static void die(const char *what) {
    std::cout << what << ": " << hspmv_last_error() << std::endl;
    std::exit(1);
  }

public:
  CSRk_Graph() : A_(NULL), N_(0), M_(0), NNZ_(0), k_(0), y_(NULL) {}

  CSRk_Graph(long nRows, long nCols, long nnz, unsigned int *rVec, unsigned int *cVec, val_type *value,
             std::string kernelCalled, std::string /*inOrderType*/, std::string inCoarsenType,
             bool /*isTuned*/, int inK, int *inSupRowSizes)
      : A_(NULL), N_(nRows), M_(nCols), NNZ_(nnz), k_(inK), y_(NULL) {
    if (kernelCalled != "SpMV" || (inK > 1 && inCoarsenType != "HAND")) {
      std::cout << "Undefined kernel invoked" << std::endl;
      std::exit(1);
    }
    for (int i = 0; i < inK - 1; i++) sizes_.push_back(inSupRowSizes[i]);
    if (hspmv_create_csr(&A_, nRows, nCols, nnz, (const int32_t *)rVec, (const int32_t *)cVec, value,
                         HSPMV_VAL_DTYPE, 0, /*keep_host=*/1))
      die("CSRk_Graph");
    bind_();
  }

  ~CSRk_Graph() { hspmv_destroy(A_); }

  void putInCSRkFormat() {
    if (k_ < 2) return;
    if (hspmv_build_csrk(A_, k_, sizes_.data())) die("putInCSRkFormat");
    perm_.resize((size_t)N_);
    if (hspmv_csrk_get(A_, perm_.data(), NULL, NULL)) die("getPermutation");
    bind_(); // the reordered arrays replaced the uploaded ones
  }

  unsigned int *getPermutation() { return perm_.empty() ? NULL : perm_.data(); }

  long getNumCoarsestRows() {
    int64_t n_sr = 0, n_ssr = 0;
    int k = 1;
    hspmv_matrix_info(A_, NULL, NULL, NULL, NULL, &k, &n_sr, &n_ssr);
    return k == 3 ? (long)n_ssr : k == 2 ? (long)n_sr : N_;
  }

  void setX(val_type *x) {
    if (hspmv_set_x(A_, x)) die("setX");
  }
  void setY(val_type *y) {
    y_ = y;
    if (hspmv_set_y(A_, y)) die("setY");
  }
  val_type *getY() { // permuted row order, as the reference's device y (csrk.cu:109-112)
    if (y_ && hspmv_get_y(A_, y_, 0)) die("getY");
    return y_;
  }

  // CUDA flavour: launch on the device-resident vectors and return (csrk.cuh:347)
  void SpMV() {
    if (hspmv_spmv(A_, HSPMV_ALGO_AUTO, NULL, NULL, NULL)) die("SpMV");
  }
  // CPU flavour: y is filled in the permuted row order (spmv-csrk/csrk.cpp:149; compared with
  // y_csrser[permutation[i]] at spmv-csrk/spmv.cpp:203)
  void SpMV(val_type *&y) {
    SpMV();
    if (hspmv_get_y(A_, y, 0)) die("SpMV");
  }

  hspmv_matrix *handle() { return A_; }
};

#endif
