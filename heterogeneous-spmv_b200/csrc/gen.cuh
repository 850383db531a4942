// Device-side generators for the stencil shapes of BASELINE.json (configs 1, 2, 5), so that the
// 1.7 G-nonzero configuration never has to exist on the host.  Same closed forms and the same
// value formulas as generators.py (stencil27 / laplace2d_5pt, natural order): the tests compare
// the two element by element on small grids.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hspmv {

// entries per grid line position t in [0,g): 3 inside, 2 at either end (1 if g == 1)
__host__ __device__ inline int64_t line_len(int64_t t, int64_t g) {
  return g == 1 ? 1 : 3 - (t == 0) - (t == g - 1);
}
// sum of line_len over positions < t
__host__ __device__ inline int64_t line_prefix(int64_t t, int64_t g) {
  return g == 1 ? t : (t == 0 ? 0 : 3 * t - 1);
}
// nonzeros in rows < i of the g^3 27-point stencil
__host__ __device__ inline int64_t stencil27_before(int64_t i, int64_t g) {
  const int64_t S = g == 1 ? 1 : 3 * g - 2;
  if (i >= g * g * g) return S * S * S;
  const int64_t x = i % g, y = (i / g) % g, z = i / (g * g);
  return line_prefix(z, g) * S * S +
         line_len(z, g) * (line_prefix(y, g) * S + line_len(y, g) * line_prefix(x, g));
}
// nonzeros in rows < i of the g^2 5-point Laplacian
__host__ __device__ inline int64_t laplace5_before(int64_t i, int64_t g) {
  const int64_t n = g * g;
  if (i >= n) i = n;
  if (g == 1) return i;
  const int64_t r = i / g, c = i % g;
  const int64_t miss_up = i < g ? i : g;
  const int64_t miss_down = i > (g - 1) * g ? i - (g - 1) * g : 0;
  const int64_t miss_left = r + (c > 0 ? 1 : 0);
  const int64_t miss_right = r;
  return 5 * i - miss_up - miss_down - miss_left - miss_right;
}

// generators.py:value_of(mode='varied')
template <typename T> __host__ __device__ inline T varied_value(int64_t row, int64_t col) {
  if (row == col) return (T)(32.0 + (double)((7 * row) % 64) / 64.0);
  return (T)(-(1.0 + (double)((31 * row + 17 * col) % 64) / 64.0));
}

template <typename T>
__global__ void stencil27_kernel(int64_t g, int64_t row_lo, int64_t row_hi, int64_t col_shift,
                                 int mode, int *__restrict__ rp, int *__restrict__ ci,
                                 T *__restrict__ val) {
  const int64_t base = stencil27_before(row_lo, g);
  for (int64_t i = row_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= row_hi;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t e = stencil27_before(i, g) - base;
    rp[i - row_lo] = (int)e;
    if (i == row_hi) break;
    const int64_t x = i % g, y = (i / g) % g, z = i / (g * g);
    for (int dz = -1; dz <= 1; ++dz) {
      if (z + dz < 0 || z + dz >= g) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        if (y + dy < 0 || y + dy >= g) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          if (x + dx < 0 || x + dx >= g) continue;
          const int64_t c = i + (dz * g + dy) * g + dx;
          ci[e] = (int)(c - col_shift);
          val[e] = mode == 0 ? (T)(c == i ? 26.0 : -1.0) : varied_value<T>(i, c);
          ++e;
        }
      }
    }
  }
}

template <typename T>
__global__ void laplace5_kernel(int64_t g, int64_t row_lo, int64_t row_hi, int64_t col_shift,
                                int mode, int *__restrict__ rp, int *__restrict__ ci,
                                T *__restrict__ val) {
  const int64_t base = laplace5_before(row_lo, g);
  for (int64_t i = row_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= row_hi;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t e = laplace5_before(i, g) - base;
    rp[i - row_lo] = (int)e;
    if (i == row_hi) break;
    const int64_t r = i / g, c = i % g;
    const int dr[5] = {-1, 0, 0, 0, 1}, dc[5] = {0, -1, 0, 1, 0};
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      if (r + dr[t] < 0 || r + dr[t] >= g || c + dc[t] < 0 || c + dc[t] >= g) continue;
      const int64_t j = i + dr[t] * g + dc[t];
      ci[e] = (int)(j - col_shift);
      val[e] = mode == 0 ? (T)(j == i ? 4.0 : -1.0) : varied_value<T>(i, j);
      ++e;
    }
  }
}

// x_i = 1 + (i mod 97)/128 for global index i = offset + local index (generators.py:vector_x)
template <typename T>
__global__ void vector_x_kernel(int64_t n, int64_t offset, T *__restrict__ x) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    x[i] = (T)(1.0 + (double)((i + offset) % 97) / 128.0);
}


// x_i = (1 + ((i + offset) mod 97)/128) * scale on up to 8 index ranges [lo, hi) of one vector, in one launch
// (hspmv_fill_x_segments: the per-step x update of the multi-GPU parity runs)
struct FillSegs {
  int n;
  int64_t lo[8], hi[8];
};
template <typename T>
__global__ void vector_x_segments_kernel(FillSegs segs, int64_t offset, double scale, T *__restrict__ x) {
  for (int s = 0; s < segs.n; ++s)
    for (int64_t i = segs.lo[s] + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < segs.hi[s];
         i += (int64_t)gridDim.x * blockDim.x)
      x[i] = (T)((1.0 + (double)((i + offset) % 97) / 128.0) * scale);
}

} // namespace hspmv
