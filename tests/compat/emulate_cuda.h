// Minimal host emulation of the CUDA execution model, enough to run the kernels of include/csrk.cuh on a machine
// without a GPU (tests/test_compat_headers.py): one std::thread per CUDA thread of a block, blocks one after the
// other, warp shuffles as a lock-step exchange between the 32 threads of a warp.  It exists to check what cannot be
// seen by compiling: indexing, the lanes-per-row grouping, and that every lane of a warp reaches every shuffle (a lane
// that exits or skips one makes the exchange read a dead lane, which is recorded as an error instead of being
// undefined behaviour as on the device).  Test infrastructure only.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3e {
  unsigned x, y, z;
};
static thread_local uint3e threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

template <typename T> static inline T __ldg(const T *p) { return *p; }

namespace emu {

struct Warp {
  std::mutex m;
  std::condition_variable cv;
  int alive = 0, arrived = 0;
  uint64_t generation = 0;
  unsigned char slot[32][8];
  bool present[32];   // lane exists and has not exited
  bool posted[32];    // lane took part in the current exchange
};
static thread_local Warp *warp = nullptr;
static thread_local int lane = 0;
static std::atomic<int> errors{0};       // shuffles that read a lane which had exited / never existed
static std::atomic<long> shuffles{0};

// all live lanes of the warp meet here; returns when everybody has posted
template <typename T> static T exchange(T v, int src) {
  Warp &w = *warp;
  std::unique_lock<std::mutex> lk(w.m);
  std::memcpy(w.slot[lane], &v, sizeof(T));
  w.posted[lane] = true;
  const uint64_t gen = w.generation;
  if (++w.arrived == w.alive) {
    // last one in: publish and release (slots are read before anybody may post again, see the second meeting)
    w.arrived = 0;
    ++w.generation;
    w.cv.notify_all();
  } else {
    w.cv.wait(lk, [&] { return w.generation != gen; });
  }
  T out = v;
  if (src < 0 || src >= 32 || !w.present[src] || !w.posted[src])
    ++errors;
  else
    std::memcpy(&out, w.slot[src], sizeof(T));
  // second meeting: nobody overwrites a slot before all have read
  const uint64_t gen2 = w.generation;
  if (++w.arrived == w.alive) {
    w.arrived = 0;
    ++w.generation;
    for (int i = 0; i < 32; ++i) w.posted[i] = false;
    w.cv.notify_all();
  } else {
    w.cv.wait(lk, [&] { return w.generation != gen2; });
  }
  ++shuffles;
  return out;
}

static void lane_exit() {
  Warp &w = *warp;
  std::unique_lock<std::mutex> lk(w.m);
  w.present[lane] = false;
  --w.alive;
  if (w.alive > 0 && w.arrived == w.alive) { // the others were waiting for me: they go on (and will flag the dead read)
    w.arrived = 0;
    ++w.generation;
    w.cv.notify_all();
  }
}

// kernel<<<grid, block>>>(args...)
template <typename K, typename... A> static void launch(K kernel, dim3 grid, dim3 block, A... args) {
  const unsigned nthreads = block.x * block.y * block.z;
  const unsigned nwarps = (nthreads + 31) / 32;
  for (unsigned b = 0; b < grid.x; ++b) {
    std::vector<Warp> warps(nwarps);
    for (unsigned wi = 0; wi < nwarps; ++wi) {
      Warp &w = warps[wi];
      for (int i = 0; i < 32; ++i) {
        w.present[i] = wi * 32 + (unsigned)i < nthreads;
        w.posted[i] = false;
        if (w.present[i]) ++w.alive;
      }
    }
    std::vector<std::thread> ts;
    ts.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t)
      ts.emplace_back([&, t, b] {
        threadIdx.x = t % block.x;
        threadIdx.y = (t / block.x) % block.y;
        threadIdx.z = t / (block.x * block.y);
        blockIdx.x = b;
        blockIdx.y = blockIdx.z = 0;
        blockDim = block;
        gridDim = grid;
        warp = &warps[t / 32];
        lane = (int)(t % 32);
        kernel(args...);
        lane_exit();
      });
    for (auto &t : ts) t.join();
  }
}

} // namespace emu

template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
  if (mask != 0xffffffffu) ++emu::errors;
  return emu::exchange<T>(v, emu::lane ^ lane_mask);
}
