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

#include "ptx.cuh"

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

// Tridiagonal x tridiagonal (the 5-point finite-difference Laplacian), the case that has to run at HBM speed:
// a block owns 256 consecutive i and sweeps a chunk of o with a sliding window of three grid lines held in
// registers, so every element of V is loaded once per thread (the i-neighbours come from L1), with the line
// two steps ahead already in flight.  Requires mi even and 16-byte aligned V / W.
template <typename T> struct V2;
template <> struct V2<double> { using type = double2; };
template <> struct V2<float> { using type = float2; };

template <typename T>
__global__ void __launch_bounds__(128) band_kronsum_tri_kernel(long long mi, long long mo, int ochunk,
                                                               const T* __restrict__ Bx, const T* __restrict__ By,
                                                               const T* __restrict__ V, T* W, T alpha, T beta) {
  using P = typename V2<T>::type;
  const long long i0 = ((long long)blockIdx.x * 128 + threadIdx.x) * 2;
  // (a thread past the edge keeps running with clamped rows so that the warp shuffles stay converged)
  const bool active = i0 < mi;
  const long long i0c = active ? i0 : mi - 2;
  const long long o_begin = (long long)blockIdx.y * ochunk;
  const long long o_end = min(mo, o_begin + (long long)ochunk);
  const long long plane = mi * mo;
  const T* v = V + (long long)blockIdx.z * plane + i0c;
  T* w = W + (long long)blockIdx.z * plane + i0c;
  // rows i0, i0+1 of By: sub-, main, super-diagonal
  const T a0 = __ldg(By + i0c), b0 = __ldg(By + i0c + mi), c0 = __ldg(By + i0c + 2 * mi);
  const T a1 = __ldg(By + i0c + 1), b1 = __ldg(By + i0c + 1 + mi), c1 = __ldg(By + i0c + 1 + 2 * mi);
  const bool has_left = i0c > 0, has_right = i0c + 2 < mi;
  const int lane = threadIdx.x & 31;
  auto line = [&](long long o) -> P {
    if (o < 0 || o >= mo) return P{T(0), T(0)};
    return __ldcs(reinterpret_cast<const P*>(v + o * mi));
  };
  // i-neighbours of the pair come from the neighbouring lanes by shuffle; only the two edge lanes of a warp
  // load a scalar (lane 0: v[i0-1], lane 31: v[i0+2]), and they do so PF lines ahead like the main loads
  auto edge = [&](long long o) -> T {
    if (o < 0 || o >= mo) return T(0);
    if (lane == 0) return has_left ? v[o * mi - 1] : T(0);
    if (lane == 31) return has_right ? v[o * mi + 2] : T(0);
    return T(0);
  };
  constexpr int PF = 4;                     // lines in flight ahead of the one being computed
  P q[PF + 2];                              // q[0] = line o-1, q[1] = line o, q[2..PF+1] = lines o+1 .. o+PF
  T e[PF + 1];                              // e[0] = edge scalar of line o, e[1..PF] of lines o+1 .. o+PF
#pragma unroll
  for (int u = 0; u < PF + 2; u++) q[u] = line(o_begin - 1 + u);
#pragma unroll
  for (int u = 0; u < PF + 1; u++) e[u] = edge(o_begin + u);
  for (long long o = o_begin; o < o_end; o++) {
    const P far = line(o + PF + 1);
    const T efar = edge(o + PF + 1);
    const P prev = q[0], cur = q[1], nxt = q[2];
    T left = __shfl_up_sync(0xffffffffu, cur.y, 1);
    T right = __shfl_down_sync(0xffffffffu, cur.x, 1);
    if (lane == 0) left = e[0];
    if (lane == 31) right = e[0];
    if (!has_left) left = T(0);      // outside the grid: never let a neighbour lane's value (or a NaN) in
    if (!has_right) right = T(0);
    const T xs = __ldg(Bx + o), xm = __ldg(Bx + o + mo), xp = __ldg(Bx + o + 2 * mo);   // row o of Bx (block-uniform)
    T r0 = fma(a0, left, fma(b0, cur.x, c0 * cur.y));
    T r1 = fma(a1, cur.x, fma(b1, cur.y, c1 * right));
    r0 = fma(xs, prev.x, fma(xm, cur.x, fma(xp, nxt.x, r0)));
    r1 = fma(xs, prev.y, fma(xm, cur.y, fma(xp, nxt.y, r1)));
    r0 *= alpha;
    r1 *= alpha;
    P* dst = reinterpret_cast<P*>(w + o * mi);
    if (beta != T(0)) {
      const P old = *dst;
      r0 = fma(beta, old.x, r0);
      r1 = fma(beta, old.y, r1);
    }
    if (active) __stcs(dst, P{r0, r1});
#pragma unroll
    for (int u = 0; u < PF + 1; u++) q[u] = q[u + 1];
    q[PF + 1] = far;
#pragma unroll
    for (int u = 0; u < PF; u++) e[u] = e[u + 1];
    e[PF] = efar;
  }
}

// TMA-tiled 5-point stencil (the version that streams at HBM speed): a persistent CTA walks tiles of
// 254 x 16 grid points; a producer warp keeps TSTAGES halo tiles (256 x 18, zero-filled outside the grid by TMA)
// in flight through a full/empty mbarrier ring, eight consumer warps read their five neighbours from shared
// memory and write W with coalesced stores.  Bytes in flight per SM come from the bulk copies (3 x 36 KB), not
// from registers, which is what the register-window version above lacked (ncu: 89 % long-scoreboard stalls).
// The innermost TMA coordinate must keep the box start 16-byte aligned (an odd FP64 coordinate is an illegal
// instruction, tools/tma_probe.cu), so the left halo is 16 bytes wide (2 doubles / 4 floats) and so is the
// unused strip on the right: interior = 256 - 2*HL points per tile row.
constexpr int TTI = 256, TTO = 18, TO_IN = TTO - 2, TSTAGES = 3, TTHREADS = 288;
template <typename T> struct Tri { static constexpr int HL = 16 / (int)sizeof(T); static constexpr int IN = TTI - 2 * HL; };
// READW (5-arg, beta != 0): the old W tile (IN x 16) rides in the same ring stage, also by TMA
template <typename T, bool READW>
struct TriStage { static constexpr int BYTES = TTI * TTO * (int)sizeof(T) + (READW ? TTI * TO_IN * (int)sizeof(T) : 0); };
template <typename T, bool READW = false>
constexpr int tri_smem_bytes() { return TSTAGES * TriStage<T, READW>::BYTES + 1024 + 64; }

struct TriParams {
  long long mi, mo, k;
  int tiles_i, tiles_o;
  long long total_tiles;   // tiles_i * tiles_o * k
};

template <typename T, bool READW>
__global__ void __launch_bounds__(TTHREADS, 1)
    band_kronsum_tri_tma_kernel(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmW,
                                const TriParams p, const T* __restrict__ Bx, const T* __restrict__ By, T* W, T alpha,
                                T beta) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 127u) & ~127u;
  const T* tiles = reinterpret_cast<const T*>(smem_raw + (smem - smem_u32(smem_raw)));
  constexpr uint32_t TILE_BYTES = TTI * TTO * sizeof(T);
  constexpr uint32_t STAGE_BYTES = TriStage<T, READW>::BYTES;
  constexpr int HL = Tri<T>::HL, TI_IN = Tri<T>::IN;
  constexpr uint32_t WTILE_BYTES = TI_IN * TO_IN * sizeof(T);   // what the W box delivers (box {IN, 16})
  const uint32_t full_bar = smem + TSTAGES * STAGE_BYTES, empty_bar = full_bar + 8 * TSTAGES;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TSTAGES; s++) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 8); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long per_col = (long long)p.tiles_i * p.tiles_o;
  if (warp == 8) {
    // ---- producer ----
    int stage = 0; uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int j = (int)(tile / per_col);
      const long long rem = tile - (long long)j * per_col;
      const int to = (int)(rem / p.tiles_i), ti = (int)(rem - (long long)to * p.tiles_i);
      mbar_wait(empty_bar + 8 * stage, phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(full_bar + 8 * stage, TILE_BYTES + (READW ? WTILE_BYTES : 0u));
        tma_load_3d(smem + stage * STAGE_BYTES, &tmV, full_bar + 8 * stage, ti * TI_IN - HL, to * TO_IN - 1, j);
        if (READW) tma_load_3d(smem + stage * STAGE_BYTES + TILE_BYTES, &tmW, full_bar + 8 * stage, ti * TI_IN, to * TO_IN, j);
      }
      __syncwarp();
      if (++stage == TSTAGES) { stage = 0; phase ^= 1; }
    }
    return;
  }
  // ---- consumers: thread t <-> tile column t (grid index i = ti*TI_IN - HL + t), interior t = HL .. HL+TI_IN-1 ----
  const int t = threadIdx.x;   // 0..255
  int stage = 0; uint32_t phase = 0;
  const long long plane = p.mi * p.mo;
  for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int j = (int)(tile / per_col);
    const long long rem = tile - (long long)j * per_col;
    const int to = (int)(rem / p.tiles_i), ti = (int)(rem - (long long)to * p.tiles_i);
    const long long i = (long long)ti * TI_IN - HL + t;
    const bool interior = (t >= HL) && (t < HL + TI_IN) && (i < p.mi);
    T a = T(0), b = T(0), c = T(0);
    if (interior) { a = __ldg(By + i); b = __ldg(By + i + p.mi); c = __ldg(By + i + 2 * p.mi); }
    const long long o0 = (long long)to * TO_IN;
    T* wcol = W + (long long)j * plane + i;
    if (!READW) {
      mbar_wait(full_bar + 8 * stage, phase);
      const T* s = reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(tiles) + (size_t)stage * STAGE_BYTES);
      if (interior) {
#pragma unroll 4
        for (int r = 1; r <= TO_IN; r++) {
          const long long o = o0 + r - 1;
          if (o >= p.mo) break;
          const T xs = __ldg(Bx + o), xm = __ldg(Bx + o + p.mo), xp = __ldg(Bx + o + 2 * p.mo);
          const T* row = s + r * TTI + t;
          T v = fma(a, row[-1], fma(b, row[0], c * row[1]));
          v = fma(xs, row[-TTI], fma(xm, row[0], fma(xp, row[TTI], v)));
          __stcs(wcol + o * p.mi, alpha * v);
        }
      }
    } else {
      // the old W values of this tile arrived with the halo tile (box {IN, 16}: row r-1, column t - HL)
      mbar_wait(full_bar + 8 * stage, phase);
      const T* s = reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(tiles) + (size_t)stage * STAGE_BYTES);
      const T* wold = reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(s) + TILE_BYTES) + (t - HL);
      if (interior) {
#pragma unroll 4
        for (int r = 1; r <= TO_IN; r++) {
          const long long o = o0 + r - 1;
          if (o >= p.mo) break;
          const T xs = __ldg(Bx + o), xm = __ldg(Bx + o + p.mo), xp = __ldg(Bx + o + 2 * p.mo);
          const T* row = s + r * TTI + t;
          T v = fma(a, row[-1], fma(b, row[0], c * row[1]));
          v = fma(xs, row[-TTI], fma(xm, row[0], fma(xp, row[TTI], v)));
          __stcs(wcol + o * p.mi, fma(beta, wold[(r - 1) * TI_IN], alpha * v));
        }
      }
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(empty_bar + 8 * stage);
    if (++stage == TSTAGES) { stage = 0; phase ^= 1; }
  }
}

}  // namespace band
}  // namespace scimlop
