// Device side of the CSR-k build (bandk_gpu.cu); the coarse-graph work stays in bandk.cpp.
#pragma once
#include "host_index.h"

#include <cuda_runtime.h>

namespace hspmv {

// returned when an input exceeds a shared-memory capacity of the device builder (or scratch
// memory is short): the caller runs the host builder instead, which has no such limits
enum { BANDK_USE_HOST = 1000 };
// longest row (and longest lower-neighbour list of a coarse vertex) the warp sorters take
enum { BANDK_MAX_ROW = 1024 };

// Level-1 coarse graph of the CSR matrix (d_rp, d_ci) in HBM for the row groups `start`
// (bandk_group_rows).  Same arrays, bit for bit, as build_coarse() in bandk.cpp.
int coarse_level1_device(int64_t n, int64_t nnz, const u32 *d_rp, const u32 *d_ci, const uvec &start,
                         int sms, cudaStream_t st, CoarseGraph &out, int64_t *launches);

// A' = P A P^T in HBM: row i of A' is row perm[i] of A with columns renumbered by the inverse
// permutation and sorted (reorder_f32/f64 in bandk.cpp).  d_rp2 has n+1 entries.
int reorder_device(int dtype, int64_t n, int64_t nnz, const u32 *d_perm, const u32 *d_rp, const u32 *d_ci,
                   const void *d_val, u32 *d_rp2, u32 *d_ci2, void *d_val2, int sms, cudaStream_t st,
                   int64_t *launches);

} // namespace hspmv
