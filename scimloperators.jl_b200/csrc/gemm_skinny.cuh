// gemm_skinny.cuh -- Y(M x k) = alpha * op(A) * X(N x k) + beta * Y for FEW columns (k <= 48 / 64, in passes of 8): the vector-input
// `mul!(w, L::MatrixOperator, v)` of an ODE / Krylov caller (src/matrix.jl:337-350 with a Vector or a thin
// matrix) and the WOperator / Jacobian products around it.  HBM-bound: the matrix is read exactly once.
//
//   * the matrix streams through a 4-stage ring of 32 KB TMA tiles (2-D `cp.async.bulk.tensor`, no swizzle
//     needed: lanes read consecutive 16-byte pieces), so the bytes in flight per SM do not depend on how many
//     threads there are; the matching slice of X rides in the same stage;
//   * NT (A stored M x N): a thread owns VEC = 16/sizeof(T) consecutive rows and all (<= 8) columns;
//     T  (A stored N x M, op = transpose): a warp owns 4 outputs, lanes stride along the contiguous reduction
//     index and are combined with shuffles at the end;
//   * when there are too few row tiles to fill 148 SMs, the reduction range is split over a thread-block
//     CLUSTER (<= 8 CTAs) and the partial results are summed through distributed shared memory in rank order
//     (deterministic; no atomics, no workspace).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace scimlop {
namespace skinny {

constexpr int CONSUMERS = 128;
constexpr int THREADS = CONSUMERS + 32;  // + one producer warp
constexpr int STAGES = 4;
constexpr int TO = 16;                   // "other" extent of a matrix tile (NT: reduction j; T: outputs i)
constexpr int MAXC = 8;                  // columns of X per pass
constexpr int A_BYTES = 32768;           // one matrix tile: (128 * VEC contiguous) x 16
constexpr int X_BYTES = 16384;           // room for the X slice of a stage (T: 128*VEC x 8; NT: 16 x 8)
constexpr int STAGE_BYTES = A_BYTES + X_BYTES;
constexpr int RED_BYTES = 16384;         // partial results of one CTA (NT: 128*VEC x 8; T: 16 x 8)
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + RED_BYTES + 128;

struct Params {
  long long M, N;        // op(A) is M x N
  int k;                 // columns of X / Y
  long long ldy;
  void* Y;
  double alpha, beta;
  int splits;            // cluster size along the reduction
  int tiles_red;         // number of reduction tiles
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
template <typename T> struct Vec;
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };
template <> struct Vec<float> { using type = float4; static constexpr int N = 4; };
__device__ __forceinline__ void unpack(const double2& v, double (&o)[2]) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ void unpack(const float4& v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }

// grid = (splits, row tiles, column passes), cluster = (splits, 1, 1)
template <typename T, bool TRANS, int KC>
__global__ void __launch_bounds__(THREADS, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmX, const Params p) {
  constexpr int VEC = Vec<T>::N;
  constexpr int TC = 128 * VEC;       // contiguous extent of a tile (NT: rows; T: reduction)
  constexpr int NB = TC / 256;        // TMA boxes per tile (a box dimension is at most 256)
  using V = typename Vec<T>::type;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (smem - smem_u32(smem_raw));
  const uint32_t red = smem + STAGES * STAGE_BYTES;
  T* red_gen = reinterpret_cast<T*>(sgen + STAGES * STAGE_BYTES);
  const uint32_t full_bar = red + RED_BYTES, empty_bar = full_bar + 8 * STAGES;
  const uint32_t scratch = empty_bar + 8 * STAGES;  // one word per consumer warp (mbar_arrive_after_loads)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, nsplit = gridDim.x;
  const int otile = blockIdx.y;                     // output tile: NT 128*VEC rows, T 16 outputs
  const int c0 = blockIdx.z * MAXC;                 // first column of this pass
  const int ntiles = (p.tiles_red - split + nsplit - 1) / nsplit;   // reduction tiles split, split+nsplit, ...

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, CONSUMERS / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  constexpr uint32_t TX = A_BYTES + (TRANS ? TC : TO) * KC * sizeof(T);
  if (warp == CONSUMERS / 32) {
    // ================= TMA producer =================
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmX);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < ntiles; it++) {
        const int rt = split + it * nsplit;
        mbar_wait(empty_bar + 8 * stage, phase ^ 1);
        const uint32_t sA = smem + stage * STAGE_BYTES, sX = sA + A_BYTES, bar = full_bar + 8 * stage;
        mbar_arrive_expect_tx(bar, TX);
        if (!TRANS) {
#pragma unroll
          for (int b = 0; b < NB; b++) tma_load_2d(sA + b * (A_BYTES / NB), &tmA, bar, otile * TC + 256 * b, rt * TO);
          tma_load_2d(sX, &tmX, bar, rt * TO, c0);
        } else {
#pragma unroll
          for (int b = 0; b < NB; b++) {
            tma_load_2d(sA + b * (A_BYTES / NB), &tmA, bar, rt * TC + 256 * b, otile * TO);
            tma_load_2d(sX + b * (256 * KC * (int)sizeof(T)), &tmX, bar, rt * TC + 256 * b, c0);
          }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ================= consumers =================
    int stage = 0;
    uint32_t phase = 0;
    if (!TRANS) {
      // thread owns rows otile*TC + VEC*tid .. +VEC-1
      T acc[KC][VEC];
#pragma unroll
      for (int c = 0; c < KC; c++)
#pragma unroll
        for (int r = 0; r < VEC; r++) acc[c][r] = T(0);
      const int cc = threadIdx.x * VEC;
      const uint32_t aoff = (uint32_t)(((cc >> 8) * TO) * 256 + (cc & 255)) * sizeof(T);
      for (int it = 0; it < ntiles; it++) {
        mbar_wait(full_bar + 8 * stage, phase);
        const uint8_t* As = sgen + stage * STAGE_BYTES + aoff;
        const T* xs = reinterpret_cast<const T*>(sgen + stage * STAGE_BYTES + A_BYTES);
        uint32_t digest = 0;
#pragma unroll
        for (int j0 = 0; j0 < TO; j0 += VEC) {
          T a[VEC][VEC];
#pragma unroll
          for (int jj = 0; jj < VEC; jj++) unpack(*reinterpret_cast<const V*>(As + (j0 + jj) * 256 * sizeof(T)), a[jj]);
#pragma unroll
          for (int c = 0; c < KC; c++) {
            T x[VEC];
            unpack(*reinterpret_cast<const V*>(xs + c * TO + j0), x);   // broadcast: every lane reads the same 16 bytes
#pragma unroll
            for (int jj = 0; jj < VEC; jj++)
#pragma unroll
              for (int r = 0; r < VEC; r++) acc[c][r] = fma(a[jj][r], x[jj], acc[c][r]);
          }
        }
        // every load of this stage feeds the accumulators: a digest of them orders the slot release after the loads
#pragma unroll
        for (int c = 0; c < KC; c++) digest ^= fold_bits(acc[c][0]);
        __syncwarp();
        if (lane == 0) mbar_arrive_after_loads(empty_bar + 8 * stage, scratch + 4 * warp, digest);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
#pragma unroll
      for (int c = 0; c < KC; c++)
#pragma unroll
        for (int r = 0; r < VEC; r++) red_gen[c * TC + cc + r] = acc[c][r];
    } else {
      // warp owns outputs otile*16 + 4*warp .. +3; lane covers reduction indices VEC*(lane + 32 q), q < 4
      T acc[4][KC];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < KC; c++) acc[i][c] = T(0);
      for (int it = 0; it < ntiles; it++) {
        mbar_wait(full_bar + 8 * stage, phase);
        const uint8_t* As = sgen + stage * STAGE_BYTES;
        const uint8_t* xs = As + A_BYTES;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int cc = (lane + 32 * q) * VEC;
          const int box = cc >> 8, in = cc & 255;
          T xv[KC][VEC];
#pragma unroll
          for (int c = 0; c < KC; c++)
            unpack(*reinterpret_cast<const V*>(xs + ((box * KC + c) * 256 + in) * sizeof(T)), xv[c]);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            T a[VEC];
            unpack(*reinterpret_cast<const V*>(As + ((box * TO + 4 * warp + i) * 256 + in) * sizeof(T)), a);
#pragma unroll
            for (int c = 0; c < KC; c++)
#pragma unroll
              for (int r = 0; r < VEC; r++) acc[i][c] = fma(a[r], xv[c][r], acc[i][c]);
          }
        }
        uint32_t digest = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int c = 0; c < KC; c++) digest ^= fold_bits(acc[i][c]);
        __syncwarp();
        if (lane == 0) mbar_arrive_after_loads(empty_bar + 8 * stage, scratch + 4 * warp, digest);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < KC; c++) {
          T v = acc[i][c];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) red_gen[c * TO + 4 * warp + i] = v;
        }
    }
  }

  // ---- combine the splits through distributed shared memory, then alpha / beta and the store.  Every rank sums
  //      (in rank order, so the result does not depend on who adds) and stores its own 1/nsplit of the outputs. ----
  cluster_sync_all();
  if (warp < CONSUMERS / 32) {
    T* Y = static_cast<T*>(p.Y);
    const T alpha = (T)p.alpha, beta = (T)p.beta;
    constexpr int ROWS = TRANS ? TO : TC;      // outputs of this CTA
    const int total = ROWS * KC;
    const int chunk = (total + nsplit - 1) / nsplit;
    const int e0 = (int)cluster_rank() * chunk, e1 = min(total, e0 + chunk);
    for (int e = e0 + threadIdx.x; e < e1; e += CONSUMERS) {
      const int c = e / ROWS, r = e - c * ROWS;
      const long long row = (long long)otile * ROWS + r;
      if (row >= p.M || c0 + c >= p.k) continue;
      T part[8];
#pragma unroll
      for (int s = 0; s < 8; s++)
        part[s] = s < nsplit ? ld_cluster<T>(map_to_rank(red + (uint32_t)e * sizeof(T), s)) : T(0);
      T v = part[0];
#pragma unroll
      for (int s = 1; s < 8; s++) v += part[s];
      T* dst = Y + row + (long long)(c0 + c) * p.ldy;
      *dst = (beta == T(0)) ? alpha * v : fma(alpha, v, beta * *dst);
    }
  }
  cluster_sync_all();   // peers stay resident until rank 0 has read their partial results
}

}  // namespace skinny
}  // namespace scimlop
