// Multi-GPU SpMV inside ONE process, behind the C ABI (include/hspmv.h, hspmv_mg_*): what SURVEY.md 8(b)
// asks the ABI to export ("multi-GPU init (device list) + partition query ... one stream + comm per device")
// so that a C/C++ driver in the style of cuda-spmv-csrk/cuda/spmv.cu:111-266 can use every GPU of the box.
// The reference has no multi-GPU code; the rules are those of dist.py (one process per GPU), restated for one
// host thread that owns all devices:
//   * rows are cut into contiguous nnz-balanced blocks (hspmv_partition_rows; CSR-k: at super-row /
//     super-super-row boundaries), each device keeps its block with row pointers rebased and column ids
//     localised to the window [cmin, cmax] of x it references (hspmv_slice_rows);
//   * banded matrices: two x buffers per device (ping-pong by step parity); per step and device, all
//     stream-ordered and without host synchronisation: main stream = ONE kernel (hspmv_spmv_halo: interior tiles,
//     in-kernel wait on the neighbours' flag words, boundary tiles), comm stream = cudaMemcpyPeerAsync of my
//     boundary slices into the neighbours' halo regions, then cuStreamWriteValue32 of the step number into their
//     flag words.  No acknowledgements: see dist.PeerHalo for the argument (it needs senders == receivers);
//   * CSR-k shards: same transport, the main stream waits for the flags with cuStreamWaitValue32, then one launch;
//   * anything else (power-law matrices whose window is all of x, unsymmetric halo plans): x replicated on every
//     device, y row-sharded, a step is the local SpMV, no exchange.
// Only the public entry points of hspmv.h are used here: a shard is an ordinary hspmv_matrix on its device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "../../include/hspmv.h"

namespace {

thread_local std::string g_mg_err;
int mg_fail(int code, const std::string &msg) {
  g_mg_err = msg;
  cudaGetLastError(); // a failed runtime call must not be found again by the next cudaGetLastError() of the library
  return code;
}
#define MGCU(call)                                                                              \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) return mg_fail(HSPMV_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
#define MGRC(call)                                                              \
  do {                                                                          \
    int rc_ = (call);                                                           \
    if (rc_) return mg_fail(rc_, std::string(#call) + ": " + hspmv_last_error()); \
  } while (0)

struct Seg {
  int peer;
  int64_t lo, hi; // global column range [lo, hi)
};

struct Shard {
  int dev = 0;
  hspmv_matrix *A = nullptr;
  int64_t lo = 0, hi = 0, cmin = 0, cmax = -1; // rows [lo,hi), referenced columns [cmin,cmax]
  void *x[2] = {nullptr, nullptr};             // x[0] = the handle's own vector, x[1] = second buffer
  int *flags = nullptr;                        // 64 ints: [0, n_senders) data flags, [32..41] halo diagnostics
  int64_t int_a = 0, int_b = 0;                // interior rows (need no halo)
  cudaStream_t main = nullptr, comm = nullptr;
  cudaEvent_t ready[2] = {nullptr, nullptr}, pushed[2] = {nullptr, nullptr}, t0 = nullptr, t1 = nullptr;
  bool pushed_valid[2] = {false, false};
  std::vector<Seg> send, recv;
  std::vector<int> senders, receivers;
  std::vector<int> slot_at; // per receiver: my index among ITS senders
};

} // namespace

struct hspmv_mg {
  int n_dev = 0, dtype = HSPMV_F64, k = 1;
  int64_t m = 0, n = 0, nnz = 0;
  bool replicated = false;
  int step = 0;
  std::vector<Shard> s;
  std::vector<int64_t> row_lo;
  std::vector<uint32_t> perm; // CSR-k: new -> old (x and y cross the boundary in the ORIGINAL numbering)
  std::vector<uint32_t> xperm; // replicated power-law matrices: hub-column relabel, x only (new column -> old column)
  std::vector<unsigned char> stage; // host staging for permuted vectors
  size_t vs() const { return dtype == HSPMV_F64 ? 8 : 4; }
};

namespace {

// dist.halo_plan: recv = pieces of my window owned by peers, send = pieces of my rows inside peers' windows
void halo_plan(const hspmv_mg *M, int r, std::vector<Seg> &recv, std::vector<Seg> &send) {
  const Shard &me = M->s[(size_t)r];
  for (int q = 0; q < M->n_dev; ++q) {
    if (q == r) continue;
    const Shard &o = M->s[(size_t)q];
    int64_t a = std::max(me.cmin, o.lo), b = std::min(me.cmax + 1, o.hi);
    if (a < b) recv.push_back({q, a, b});
    a = std::max(o.cmin, me.lo);
    b = std::min(o.cmax + 1, me.hi);
    if (a < b) send.push_back({q, a, b});
  }
}

int destroy_shards(hspmv_mg *M) {
  for (Shard &sh : M->s) {
    cudaSetDevice(sh.dev);
    cudaDeviceSynchronize();
    if (sh.A) hspmv_destroy(sh.A);
    if (sh.x[1]) cudaFree(sh.x[1]);
    if (sh.flags) cudaFree(sh.flags);
    for (cudaEvent_t e : {sh.ready[0], sh.ready[1], sh.pushed[0], sh.pushed[1], sh.t0, sh.t1})
      if (e) cudaEventDestroy(e);
    if (sh.main) cudaStreamDestroy(sh.main);
    if (sh.comm) cudaStreamDestroy(sh.comm);
  }
  M->s.clear();
  return HSPMV_OK;
}

// `full` lives on devices[0]; cut it into shards, one per device.  Takes ownership of `full`: it is destroyed (and the
// caller's pointer cleared) once the shards exist; on an earlier failure the caller still holds it and frees it.
int build_shards(hspmv_mg *M, hspmv_matrix *&full, const int *devices, int n_dev) {
  int k = 1;
  int64_t n_sr = 0, n_ssr = 0;
  MGRC(hspmv_matrix_info(full, &M->m, &M->n, &M->nnz, &M->dtype, &k, &n_sr, &n_ssr));
  M->k = k;
  M->n_dev = n_dev;
  // ---- partition (host index work on row_ptr)
  std::vector<int32_t> rp((size_t)M->m + 1);
  MGRC(hspmv_download_csr(full, rp.data(), nullptr, nullptr));
  M->row_lo.assign((size_t)n_dev + 1, 0);
  if (k >= 2) {
    std::vector<uint32_t> sr((size_t)n_sr + 1), ssr((size_t)n_ssr + 1);
    M->perm.resize((size_t)M->m);
    int rc = hspmv_csrk_get(full, M->perm.data(), sr.data(), k == 3 ? ssr.data() : nullptr);
    if (rc == HSPMV_ESTATE) { // came from a .csr3 file: no permutation, vectors are in the file's numbering
      M->perm.clear();
      rc = hspmv_csrk_get(full, nullptr, sr.data(), k == 3 ? ssr.data() : nullptr);
    }
    if (rc) return mg_fail(rc, hspmv_last_error());
    std::vector<uint32_t> align;
    if (k == 2) {
      align = sr;
    } else {
      align.resize(ssr.size());
      for (size_t i = 0; i < ssr.size(); ++i) align[i] = sr[ssr[i]];
    }
    MGRC(hspmv_partition_rows(M->m, rp.data(), n_dev, align.data(), (int64_t)align.size() - 1, M->row_lo.data()));
  } else {
    MGRC(hspmv_partition_rows(M->m, rp.data(), n_dev, nullptr, 0, M->row_lo.data()));
  }
  std::vector<int32_t>().swap(rp);
  // ---- column windows
  M->s.assign((size_t)n_dev, Shard());
  for (int d = 0; d < n_dev; ++d) {
    Shard &sh = M->s[(size_t)d];
    sh.dev = devices[d];
    sh.lo = M->row_lo[(size_t)d];
    sh.hi = M->row_lo[(size_t)d + 1];
    MGCU(cudaSetDevice(devices[0]));
    hspmv_matrix *wide = nullptr;
    MGRC(hspmv_slice_rows(full, sh.lo, sh.hi, 0, 0, &wide));
    int rc = hspmv_col_range_device(wide, &sh.cmin, &sh.cmax);
    hspmv_destroy(wide);
    if (rc) return mg_fail(rc, hspmv_last_error());
    if (sh.cmax < sh.cmin) { // empty block
      sh.cmin = sh.lo < M->n ? sh.lo : 0;
      sh.cmax = sh.cmin;
    }
  }
  // ---- banded or replicated?  (square matrix, every window a modest part of x, symmetric exchange)
  M->replicated = n_dev > 1 && M->m != M->n;
  if (n_dev > 1 && !M->replicated) {
    for (int d = 0; d < n_dev; ++d) {
      const Shard &sh = M->s[(size_t)d];
      if ((sh.cmax - sh.cmin + 1) > (sh.hi - sh.lo) + M->n / 2) M->replicated = true;
      std::vector<Seg> recv, send;
      halo_plan(M, d, recv, send);
      std::set<int> a, b;
      for (const Seg &g : recv) a.insert(g.peer);
      for (const Seg &g : send) b.insert(g.peer);
      if (a != b || a.size() > 32) M->replicated = true;
    }
  }
  if (M->replicated && k == 1 && M->m == M->n) { // power law: columns renumbered by popularity before the cut (x only)
    M->xperm.resize((size_t)M->n);
    MGCU(cudaSetDevice(devices[0]));
    MGRC(hspmv_relabel_hub_columns(full, M->xperm.data()));
  }
  if (M->replicated || n_dev == 1) // the device holds all of x
    for (Shard &sh : M->s) {
      sh.cmin = 0;
      sh.cmax = M->n - 1;
    }
  // ---- shards
  for (int d = 0; d < n_dev; ++d) {
    Shard &sh = M->s[(size_t)d];
    MGCU(cudaSetDevice(sh.dev));
    MGRC(hspmv_slice_rows(full, sh.lo, sh.hi, sh.cmin, sh.cmax - sh.cmin + 1, &sh.A));
    MGCU(cudaStreamCreateWithFlags(&sh.main, cudaStreamNonBlocking));
    MGCU(cudaStreamCreateWithFlags(&sh.comm, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
      MGCU(cudaEventCreateWithFlags(&sh.ready[b], cudaEventDisableTiming));
      MGCU(cudaEventCreateWithFlags(&sh.pushed[b], cudaEventDisableTiming));
    }
    MGCU(cudaEventCreate(&sh.t0));
    MGCU(cudaEventCreate(&sh.t1));
    sh.x[0] = hspmv_x_device(sh.A);
    const size_t bytes = (size_t)(sh.cmax - sh.cmin + 1 + 16) * M->vs();
    MGCU(cudaMalloc(&sh.x[1], bytes));
    MGCU(cudaMemset(sh.x[1], 0, bytes));
    MGCU(cudaMalloc((void **)&sh.flags, 64 * sizeof(int)));
    MGCU(cudaMemset(sh.flags, 0, 64 * sizeof(int)));
    if (!M->replicated && n_dev > 1 && k == 1)
      MGRC(hspmv_interior_rows(sh.A, sh.lo - sh.cmin, sh.hi - sh.cmin, &sh.int_a, &sh.int_b));
  }
  MGCU(cudaSetDevice(devices[0]));
  hspmv_destroy(full);
  full = nullptr;
  if (M->replicated || n_dev == 1) return HSPMV_OK;
  // ---- exchange plan
  for (int d = 0; d < n_dev; ++d) {
    Shard &sh = M->s[(size_t)d];
    halo_plan(M, d, sh.recv, sh.send);
    std::set<int> a, b;
    for (const Seg &g : sh.recv) a.insert(g.peer);
    for (const Seg &g : sh.send) b.insert(g.peer);
    sh.senders.assign(a.begin(), a.end());
    sh.receivers.assign(b.begin(), b.end());
  }
  for (int d = 0; d < n_dev; ++d) {
    Shard &sh = M->s[(size_t)d];
    for (int q : sh.receivers) {
      const std::vector<int> &theirs = M->s[(size_t)q].senders;
      sh.slot_at.push_back((int)(std::find(theirs.begin(), theirs.end(), d) - theirs.begin()));
    }
  }
  // One untimed launch of the kernel every step will use, without a flag wait (x is all zeros): the module is loaded
  // and its shared-memory attribute set before any step has work pending on the device.
  for (Shard &sh : M->s) {
    MGCU(cudaSetDevice(sh.dev));
    if (k == 1) MGRC(hspmv_spmv_halo(sh.A, sh.int_a, sh.int_b, sh.x[0], nullptr, sh.flags, 0, 0, sh.flags + 32, sh.main));
    else MGRC(hspmv_spmv(sh.A, HSPMV_ALGO_CSRK, sh.x[0], nullptr, sh.main));
    MGCU(cudaStreamSynchronize(sh.main));
  }
  return HSPMV_OK;
}

int enable_peers(const int *devices, int n_dev) {
  for (int a = 0; a < n_dev; ++a) {
    MGCU(cudaSetDevice(devices[a]));
    for (int b = 0; b < n_dev; ++b) {
      if (a == b) continue;
      int can = 0;
      MGCU(cudaDeviceCanAccessPeer(&can, devices[a], devices[b]));
      if (!can) return mg_fail(HSPMV_EUNSUPPORTED, "devices " + std::to_string(devices[a]) + " and " + std::to_string(devices[b]) + " have no peer access");
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return mg_fail(HSPMV_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      cudaGetLastError();
    }
  }
  return HSPMV_OK;
}

int check_devices(const int *devices, int n_dev) {
  if (!devices || n_dev < 1 || n_dev > 64) return mg_fail(HSPMV_EINVAL, "bad device list");
  int count = 0;
  MGCU(cudaGetDeviceCount(&count));
  for (int i = 0; i < n_dev; ++i) {
    if (devices[i] < 0 || devices[i] >= count) return mg_fail(HSPMV_EINVAL, "device " + std::to_string(devices[i]) + " does not exist");
    for (int j = 0; j < i; ++j)
      if (devices[j] == devices[i]) return mg_fail(HSPMV_EINVAL, "device listed twice");
  }
  return n_dev > 1 ? enable_peers(devices, n_dev) : HSPMV_OK;
}

int finish_create(hspmv_mg **out, hspmv_matrix *full, const int *devices, int n_dev) {
  hspmv_mg *M = new hspmv_mg();
  int rc = build_shards(M, full, devices, n_dev);
  if (rc) {
    const std::string why = g_mg_err; // the clean-up below must not replace the reason
    destroy_shards(M);
    delete M;
    if (full) {
      cudaSetDevice(devices[0]);
      hspmv_destroy(full);
    }
    g_mg_err = why;
    return rc;
  }
  *out = M;
  return HSPMV_OK;
}

} // namespace

extern "C" {

const char *hspmv_mg_last_error(void) { return g_mg_err.c_str(); }

int hspmv_mg_create_csr(hspmv_mg **out, const int *devices, int n_dev, int64_t m, int64_t n, int64_t nnz,
                        const int32_t *row_ptr, const int32_t *col_idx, const void *val, int dtype, int index_base, int k,
                        const int *sup_row_sizes) {
  if (!out) return mg_fail(HSPMV_EINVAL, "null");
  if (k < 1 || k > 3 || (k > 1 && !sup_row_sizes)) return mg_fail(HSPMV_EINVAL, "k must be 1, 2 or 3 (with sizes)");
  int rc = check_devices(devices, n_dev);
  if (rc) return rc;
  MGCU(cudaSetDevice(devices[0]));
  hspmv_matrix *full = nullptr;
  MGRC(hspmv_create_csr(&full, m, n, nnz, row_ptr, col_idx, val, dtype, index_base, 0));
  if (k > 1 && (rc = hspmv_build_csrk(full, k, sup_row_sizes))) {
    hspmv_destroy(full);
    return mg_fail(rc, hspmv_last_error());
  }
  return finish_create(out, full, devices, n_dev);
}

int hspmv_mg_create_stencil(hspmv_mg **out, const int *devices, int n_dev, int kind, int64_t g, int dtype, int mode, int k,
                            const int *sup_row_sizes) {
  if (!out) return mg_fail(HSPMV_EINVAL, "null");
  if (k < 1 || k > 3 || (k > 1 && !sup_row_sizes)) return mg_fail(HSPMV_EINVAL, "k must be 1, 2 or 3 (with sizes)");
  int rc = check_devices(devices, n_dev);
  if (rc) return rc;
  MGCU(cudaSetDevice(devices[0]));
  hspmv_matrix *full = nullptr;
  MGRC(hspmv_create_stencil(&full, kind, g, dtype, mode, 0, -1, 0, 0));
  if (k > 1 && (rc = hspmv_build_csrk(full, k, sup_row_sizes))) {
    hspmv_destroy(full);
    return mg_fail(rc, hspmv_last_error());
  }
  return finish_create(out, full, devices, n_dev);
}

int hspmv_mg_destroy(hspmv_mg *M) {
  if (!M) return HSPMV_OK;
  destroy_shards(M);
  delete M;
  return HSPMV_OK;
}

int hspmv_mg_info(const hspmv_mg *M, int *n_dev, int64_t *m, int64_t *n, int64_t *nnz, int *dtype, int *k, int *replicated) {
  if (!M) return mg_fail(HSPMV_EINVAL, "null handle");
  if (n_dev) *n_dev = M->n_dev;
  if (m) *m = M->m;
  if (n) *n = M->n;
  if (nnz) *nnz = M->nnz;
  if (dtype) *dtype = M->dtype;
  if (k) *k = M->k;
  if (replicated) *replicated = M->replicated ? 1 : 0;
  return HSPMV_OK;
}

int hspmv_mg_partition_query(const hspmv_mg *M, int *devices, int64_t *row_lo, int64_t *col_min, int64_t *col_max,
                             int64_t *nnz_per_dev, int64_t *halo_elems) {
  if (!M) return mg_fail(HSPMV_EINVAL, "null handle");
  for (int d = 0; d < M->n_dev; ++d) {
    const Shard &sh = M->s[(size_t)d];
    if (devices) devices[d] = sh.dev;
    if (col_min) col_min[d] = sh.cmin;
    if (col_max) col_max[d] = sh.cmax;
    if (nnz_per_dev) MGRC(hspmv_matrix_info(sh.A, nullptr, nullptr, &nnz_per_dev[d], nullptr, nullptr, nullptr, nullptr));
    if (halo_elems) {
      halo_elems[d] = 0;
      for (const Seg &g : sh.recv) halo_elems[d] += g.hi - g.lo;
    }
  }
  if (row_lo)
    for (int d = 0; d <= M->n_dev; ++d) row_lo[d] = M->row_lo[(size_t)d];
  return HSPMV_OK;
}

int hspmv_mg_get_perm(const hspmv_mg *M, uint32_t *perm) {
  if (!M || !perm) return mg_fail(HSPMV_EINVAL, "null");
  if (M->perm.empty()) return mg_fail(HSPMV_ESTATE, "no permutation (plain CSR, or a .csr3 matrix)");
  std::memcpy(perm, M->perm.data(), M->perm.size() * 4);
  return HSPMV_OK;
}

void *hspmv_mg_shard(hspmv_mg *M, int d) { return (M && d >= 0 && d < M->n_dev) ? (void *)M->s[(size_t)d].A : nullptr; }

// x in the ORIGINAL numbering: every device receives the slice it OWNS (its rows' entries), in both buffers; the
// halo entries travel device to device inside hspmv_mg_spmv.  Replicated mode: every device receives all of x.
int hspmv_mg_set_x(hspmv_mg *M, const void *x_host) {
  if (!M || !x_host) return mg_fail(HSPMV_EINVAL, "null");
  const size_t vs = M->vs();
  const char *src = (const char *)x_host;
  const std::vector<uint32_t> &xp = !M->perm.empty() ? M->perm : M->xperm;
  if (!xp.empty()) { // setX: x_k[i] = x[perm[i]] (spmv-csrk/csrk.h:327-335); hub relabel: the same, for x only
    M->stage.resize((size_t)M->n * vs);
    if (vs == 8) {
      const double *x = (const double *)x_host;
      double *o = (double *)M->stage.data();
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < M->n; ++i) o[i] = x[xp[(size_t)i]];
    } else {
      const float *x = (const float *)x_host;
      float *o = (float *)M->stage.data();
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < M->n; ++i) o[i] = x[xp[(size_t)i]];
    }
    src = (const char *)M->stage.data();
  }
  for (Shard &sh : M->s) {
    MGCU(cudaSetDevice(sh.dev));
    for (int b = 0; b < 2; ++b)
      if (sh.pushed_valid[b]) MGCU(cudaStreamWaitEvent(sh.main, sh.pushed[b], 0)); // my pushes still read these slices
    if (M->replicated || M->n_dev == 1) {
      for (int b = 0; b < 2; ++b)
        MGCU(cudaMemcpyAsync(sh.x[b], src, (size_t)M->n * vs, cudaMemcpyHostToDevice, sh.main));
    } else {
      const size_t off = (size_t)(sh.lo - sh.cmin) * vs, bytes = (size_t)(sh.hi - sh.lo) * vs;
      MGCU(cudaMemcpyAsync((char *)sh.x[0] + off, src + (size_t)sh.lo * vs, bytes, cudaMemcpyHostToDevice, sh.main));
      MGCU(cudaMemcpyAsync((char *)sh.x[1] + off, (char *)sh.x[0] + off, bytes, cudaMemcpyDeviceToDevice, sh.main));
    }
  }
  for (Shard &sh : M->s) {
    MGCU(cudaSetDevice(sh.dev));
    MGCU(cudaStreamSynchronize(sh.main));
  }
  return HSPMV_OK;
}

// test hook: NaN into the halo regions of both x buffers of every device (nothing but a push may fill them)
int hspmv_mg_poison_halos(hspmv_mg *M) {
  if (!M) return mg_fail(HSPMV_EINVAL, "null handle");
  const size_t vs = M->vs();
  for (Shard &sh : M->s) {
    MGCU(cudaSetDevice(sh.dev));
    MGCU(cudaDeviceSynchronize());
    for (const Seg &g : sh.recv)
      for (int b = 0; b < 2; ++b)
        MGCU(cudaMemset((char *)sh.x[b] + (size_t)(g.lo - sh.cmin) * vs, 0xff, (size_t)(g.hi - g.lo) * vs)); // all-ones = NaN
  }
  return HSPMV_OK;
}

// One step on every device: halo exchange of the current x + local SpMV, all enqueued, nothing waited for.
int hspmv_mg_spmv(hspmv_mg *M) {
  if (!M) return mg_fail(HSPMV_EINVAL, "null handle");
  const int t = ++M->step;
  const int b = t & 1;
  const size_t vs = M->vs();
  const bool exchange = !M->replicated && M->n_dev > 1;
  if (!exchange) {
    for (Shard &sh : M->s) {
      MGCU(cudaSetDevice(sh.dev));
      MGRC(hspmv_spmv(sh.A, HSPMV_ALGO_AUTO, sh.x[b], nullptr, sh.main));
    }
    return HSPMV_OK;
  }
  // 1. every device's pushes are enqueued BEFORE any kernel: one host thread feeds all devices, and a launch may block
  //    the host (the first launch of a kernel loads its module, which waits for the device's pending work -- e.g. a
  //    cuStreamWaitValue32 that only the next device's push can satisfy); with the pushes already in their queues
  //    nothing a device waits for depends on the host any more
  for (Shard &sh : M->s) {
    MGCU(cudaSetDevice(sh.dev));
    MGCU(cudaEventRecord(sh.ready[b], sh.main)); // x of step t is complete at this point of the main stream
    // comm stream: my boundary slices into the neighbours' x[b], then their flag words
    MGCU(cudaStreamWaitEvent(sh.comm, sh.ready[b], 0));
    for (const Seg &g : sh.send) {
      Shard &q = M->s[(size_t)g.peer];
      // (UVA + peer access: a plain stream-ordered copy-engine copy, the same call the one-process-per-GPU layer uses)
      MGCU(cudaMemcpyAsync((char *)q.x[b] + (size_t)(g.lo - q.cmin) * vs, (const char *)sh.x[b] + (size_t)(g.lo - sh.cmin) * vs,
                           (size_t)(g.hi - g.lo) * vs, cudaMemcpyDefault, sh.comm));
    }
    for (size_t i = 0; i < sh.receivers.size(); ++i)
      MGRC(hspmv_stream_write(sh.comm, M->s[(size_t)sh.receivers[i]].flags + sh.slot_at[i], t));
    MGCU(cudaEventRecord(sh.pushed[b], sh.comm));
    sh.pushed_valid[b] = true;
  }
  // 2. the local products
  for (Shard &sh : M->s) {
    MGCU(cudaSetDevice(sh.dev));
    if (M->k == 1) { // one launch: interior tiles, in-kernel wait for the neighbours' flags, boundary tiles
      MGRC(hspmv_spmv_halo(sh.A, sh.int_a, sh.int_b, sh.x[b], nullptr, sh.flags, (int)sh.senders.size(), t, sh.flags + 32, sh.main));
    } else {         // CSR-k walks whole super-rows: front-end wait for the flags, then one launch
      for (size_t i = 0; i < sh.senders.size(); ++i) MGRC(hspmv_stream_wait_ge(sh.main, sh.flags + i, t));
      MGRC(hspmv_spmv(sh.A, HSPMV_ALGO_CSRK, sh.x[b], nullptr, sh.main));
    }
  }
  return HSPMV_OK;
}

int hspmv_mg_sync(hspmv_mg *M) {
  if (!M) return mg_fail(HSPMV_EINVAL, "null handle");
  for (Shard &sh : M->s) {
    MGCU(cudaSetDevice(sh.dev));
    MGCU(cudaStreamSynchronize(sh.main));
    MGCU(cudaStreamSynchronize(sh.comm));
  }
  for (Shard &sh : M->s) {
    if (M->replicated || M->n_dev == 1) break;
    int to = 0;
    MGCU(cudaSetDevice(sh.dev));
    MGCU(cudaMemcpy(&to, sh.flags + 32, 4, cudaMemcpyDeviceToHost));
    if (to) return mg_fail(HSPMV_ECUDA, "a halo flag wait timed out on device " + std::to_string(sh.dev));
  }
  return HSPMV_OK;
}

// y in the ORIGINAL row order (CSR-k: y[perm[i]] = y_k[i], like getY + the driver's check at spmv-csrk/spmv.cpp:203)
int hspmv_mg_get_y(hspmv_mg *M, void *y_host) {
  if (!M || !y_host) return mg_fail(HSPMV_EINVAL, "null");
  int rc = hspmv_mg_sync(M);
  if (rc) return rc;
  const size_t vs = M->vs();
  char *dst = (char *)y_host;
  if (!M->perm.empty()) {
    M->stage.resize((size_t)M->m * vs);
    dst = (char *)M->stage.data();
  }
  for (Shard &sh : M->s) {
    MGCU(cudaSetDevice(sh.dev));
    if (sh.hi > sh.lo)
      MGCU(cudaMemcpyAsync(dst + (size_t)sh.lo * vs, hspmv_y_device(sh.A), (size_t)(sh.hi - sh.lo) * vs, cudaMemcpyDeviceToHost, sh.main));
  }
  for (Shard &sh : M->s) {
    MGCU(cudaSetDevice(sh.dev));
    MGCU(cudaStreamSynchronize(sh.main));
  }
  if (!M->perm.empty()) {
    if (vs == 8) {
      const double *s = (const double *)M->stage.data();
      double *y = (double *)y_host;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < M->m; ++i) y[M->perm[(size_t)i]] = s[i];
    } else {
      const float *s = (const float *)M->stage.data();
      float *y = (float *)y_host;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < M->m; ++i) y[M->perm[(size_t)i]] = s[i];
    }
  }
  return HSPMV_OK;
}

// warmup + iters steps; ms_per_step[i] = longest device time of step i (events on every main stream), so that the
// drivers can print the reference's TimeMin / TimeMax / TimeAvg lines (cuda-spmv-csr/spmv.cu:240-260)
int hspmv_mg_spmv_timed(hspmv_mg *M, int warmup, int iters, float *ms_per_step) {
  if (!M || iters < 1 || !ms_per_step) return mg_fail(HSPMV_EINVAL, "bad arguments");
  for (int i = 0; i < warmup; ++i) MGRC(hspmv_mg_spmv(M));
  int rc = hspmv_mg_sync(M);
  if (rc) return rc;
  for (int i = 0; i < iters; ++i) {
    for (Shard &sh : M->s) {
      MGCU(cudaSetDevice(sh.dev));
      MGCU(cudaEventRecord(sh.t0, sh.main));
    }
    MGRC(hspmv_mg_spmv(M));
    for (Shard &sh : M->s) {
      MGCU(cudaSetDevice(sh.dev));
      MGCU(cudaEventRecord(sh.t1, sh.main));
    }
    if ((rc = hspmv_mg_sync(M))) return rc;
    float worst = 0;
    for (Shard &sh : M->s) {
      float ms = 0;
      MGCU(cudaSetDevice(sh.dev));
      MGCU(cudaEventElapsedTime(&ms, sh.t0, sh.t1));
      worst = std::max(worst, ms);
    }
    ms_per_step[i] = worst;
  }
  return HSPMV_OK;
}

// `iters` steps back to back, one pair of events per device, no host synchronisation in between: ms per step
int hspmv_mg_spmv_timed_batch(hspmv_mg *M, int warmup, int iters, float *ms_per_step) {
  if (!M || iters < 1 || !ms_per_step) return mg_fail(HSPMV_EINVAL, "bad arguments");
  for (int i = 0; i < warmup; ++i) MGRC(hspmv_mg_spmv(M));
  int rc = hspmv_mg_sync(M);
  if (rc) return rc;
  for (Shard &sh : M->s) {
    MGCU(cudaSetDevice(sh.dev));
    MGCU(cudaEventRecord(sh.t0, sh.main));
  }
  for (int i = 0; i < iters; ++i) MGRC(hspmv_mg_spmv(M));
  for (Shard &sh : M->s) {
    MGCU(cudaSetDevice(sh.dev));
    MGCU(cudaEventRecord(sh.t1, sh.main));
  }
  if ((rc = hspmv_mg_sync(M))) return rc;
  float worst = 0;
  for (Shard &sh : M->s) {
    float ms = 0;
    MGCU(cudaSetDevice(sh.dev));
    MGCU(cudaEventElapsedTime(&ms, sh.t0, sh.t1));
    worst = std::max(worst, ms);
  }
  *ms_per_step = worst / (float)iters;
  return HSPMV_OK;
}

} // extern "C"
