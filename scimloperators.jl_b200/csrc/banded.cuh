// banded.cuh -- banded (tridiagonal, pentadiagonal, ...) Kronecker factors: the finite-difference flavour of
// config 5.  `MatrixOperator` accepts any AbstractMatrix (src/matrix.jl:116), and the
// reference's own time-dependent tests use banded/sparse factors (test/basic.jl:405-426); a mode product with a
// banded matrix is a 1-D stencil along that mode, so the whole `kron(Dxx, I) + kron(I, Dyy)` matvec is ONE
// HBM-bound pass (read V once, write W once) instead of two dense GEMMs.
//
// Storage: `bands` is n x (2b+1) column-major, diagonal d (-b..b) in column d+b, aligned by ROW:
//   bands[i + n*(d+b)] = A[i, i+d]   (0 where i+d falls outside the matrix).
#pragma once
#include <cuda_runtime.h>

namespace scimlop {
namespace band {

constexpr int THREADS = 256, MAXB = 4;

// out(L*n x R) = alpha * (mode product with the banded matrix along the middle index) + beta * out
//   idx = l + L*i + L*n*r ;  out[idx] = alpha * sum_d bands[i, d+b] * in[idx + d*L] + beta * out[idx]
template <typename T>
__global__ void __launch_bounds__(THREADS) band_mode_kernel(long long total, long long L, long long n, int b,
                                                            const T* __restrict__ bands, const T* __restrict__ in,
                                                            T* out, T alpha, T beta) {
  const long long stride = (long long)gridDim.x * THREADS;
  for (long long idx = (long long)blockIdx.x * THREADS + threadIdx.x; idx < total; idx += stride) {
    const long long i = (idx / L) % n;
    T acc = T(0);
    for (int d = -b; d <= b; d++) {
      const long long ii = i + d;
      if (ii >= 0 && ii < n) acc = fma(__ldg(bands + i + n * (d + b)), __ldcs(in + idx + d * L), acc);
    }
    T o = alpha * acc;
    if (beta != T(0)) o += beta * out[idx];
    __stcs(out + idx, o);
  }
}

// Kronecker sum of two banded matrices on the (mi x mo) grid of every column:
//   w[i,o] = alpha * ( sum_d By[i,d] v[i+d,o] + sum_e Bx[o,e] v[i,o+e] ) + beta * w[i,o]
// One thread per grid point along i (contiguous), 2-D blocks so the o-neighbours hit L1/L2.
template <typename T>
__global__ void __launch_bounds__(THREADS) band_kronsum_kernel(long long mi, long long mo, long long k, int bx, int by,
                                                               const T* __restrict__ Bx, const T* __restrict__ By,
                                                               const T* __restrict__ V, T* W, T alpha, T beta) {
  // block = 64 (i) x 4 (o); grid.x over i tiles, grid.y over o tiles, grid.z over columns (looped)
  const long long i = (long long)blockIdx.x * 64 + (threadIdx.x & 63);
  const long long o = (long long)blockIdx.y * 4 + (threadIdx.x >> 6);
  if (i >= mi || o >= mo) return;
  // coefficients of this grid point's row of By / Bx, and which neighbours exist (bit dd <-> offset dd - MAXB)
  T cy[2 * MAXB + 1], cx[2 * MAXB + 1];
  unsigned my = 0, mx = 0;
#pragma unroll
  for (int dd = 0; dd < 2 * MAXB + 1; dd++) {
    const int d = dd - MAXB;
    cy[dd] = T(0);
    cx[dd] = T(0);
    if (d >= -by && d <= by && i + d >= 0 && i + d < mi) { cy[dd] = __ldg(By + i + mi * (d + by)); my |= 1u << dd; }
    if (d >= -bx && d <= bx && o + d >= 0 && o + d < mo) { cx[dd] = __ldg(Bx + o + mo * (d + bx)); mx |= 1u << dd; }
  }
  const long long plane = mi * mo;
  for (long long j = blockIdx.z; j < k; j += gridDim.z) {
    const T* v = V + j * plane + i + mi * o;
    T acc = T(0);
#pragma unroll
    for (int dd = 0; dd < 2 * MAXB + 1; dd++) {
      const int d = dd - MAXB;
      if ((my >> dd) & 1u) acc = fma(cy[dd], v[d], acc);
      if ((mx >> dd) & 1u) acc = fma(cx[dd], v[(long long)d * mi], acc);
    }
    T* w = W + j * plane + i + mi * o;
    T r = alpha * acc;
    if (beta != T(0)) r += beta * *w;
    *w = r;
  }
}

}  // namespace band
}  // namespace scimlop
