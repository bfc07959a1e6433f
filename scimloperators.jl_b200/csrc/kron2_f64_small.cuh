// kron2_f64_small.cuh -- both mode products of a two-factor Float64 Kronecker product in ONE launch, for SMALL problems
// (config 1: 100 x 100 (x) 100 x 100 on a 10^4 x 100 batch, benchmarks/tensor.jl:17-24).
//
// The general path runs a mode product as a persistent 128 x 128-tile DMMA GEMM: 95 % of the FP64 tensor peak when there are
// many waves of tiles, but config 1 has 79 and 100 tiles for 148 SMs, padded from 100 to 128 in two dimensions, in two
// launches: 55 us of kernel time for 11 us of arithmetic.  Here a CTA takes one slab X_s (n1 x n2, one batch column) into
// shared memory and computes  Z_s = F1op X_s F2op^T  with the intermediate Y = F1op X_s staying on chip (it overwrites X_s
// in shared memory once every warp holds its part of Y in accumulator registers), as the reference's fused nest does for
// one stage (ext/SciMLOperatorsLoopVectorizationExt.jl:17-23).  Tiles are 8 x 8 DMMA atoms, so the padding is to multiples
// of 8 / 4 (100 -> 104 / 100), not to 128.
//
//   stage 1:  Y(c', b)  = sum_c F1op(c', c) X(c, b)     warp w owns the 8-row strip c' = 8w ..; A = F1 (global, L1 / L2
//                                                       resident, prefetched one k-step ahead), B = X (shared memory)
//   stage 2:  Z(c', b') = sum_b Y(c', b) F2op(b', b)    warp w owns the 8-column strip b' = 8w ..; A = Y (shared
//                                                       memory), B = F2 (global, prefetched)
//
// Shared-memory pitches are = 4 (mod 16) doubles, which makes both fragment patterns (row g + pitch * t and t + pitch * g
// over a half warp) conflict-free.  mma.sync.m8n8k4.f64: lane 4g + t supplies A[g][t], B[t][g] and owns D[g][2t], D[g][2t+1].
#pragma once
#include "ptx.cuh"

namespace scimlop {
namespace k2s {

constexpr int THREADS = 512;      // 16 warps: one 8-wide strip each (MAXD / 8 = 16), four warps per sub-partition hide LDS / LDG latency
constexpr int MAXD = 128;          // factor extents supported
constexpr int MAXT = MAXD / 8;     // 8-wide strips per dimension

struct Params {
  int n1, n2, m1, m2;              // F1op: m1 x n1 (fastest mode), F2op: m2 x n2
  long long slabs;
  const double* F1; long long f1_rs, f1_cs;   // F1op(c', c) at c' * f1_rs + c * f1_cs
  const double* F2; long long f2_rs, f2_cs;
  const double* V;                 // slab s at V + s * n1 * n2
  double* W;                       // slab s at W + s * m1 * m2
  double alpha, beta;
};

// One fixed pitch for X and Y: 132 doubles (>= MAXD, = 4 mod 16), so that every fragment address is base + immediate and
// the inner loops are LDS + DMMA only (with a run-time pitch the address arithmetic cost 13 instructions per DMMA: ncu).
constexpr int PITCH = MAXD + 4;
inline int pitch_for(int) { return PITCH; }
inline size_t smem_bytes(int, int n2, int, int) { return (size_t)PITCH * ((n2 + 7) / 8 * 8) * 8; }

// NT column strips of stage 1 / MT row strips of stage 2 as a compile-time count: fully unrolled, no predicates
// Work split over the 16 warps.  A warp owns one 8-wide strip of the stage's "own" dimension and `cnt` strips of the
// other one, starting at `first`.  With 13 or 14 own strips (factor extents 97..112: config 1) one strip per warp
// would leave warps 13..15 idle AND give scheduler sub-partition 0 (warps 0, 4, 8, 12) four strips against three for
// the others -- the stage then runs at the pace of that sub-partition.  The strips beyond the twelfth are therefore cut
// along the other dimension among warps 12..15 (13 strips of 13: 4 + 3 + 3 + 3), which puts 43 / 42 / 42 / 42 tiles
// on the four sub-partitions instead of 52 / 39 / 39 / 39.
__device__ __forceinline__ void assign_strips(int warp, int own, int other, int& strip, int& first, int& cnt) {
  strip = -1; first = 0; cnt = 0;
  const int tail = own - 12;
  if (tail == 1 || tail == 2) {
    if (warp < 12) { strip = warp; cnt = other; return; }
    const int j = warp - 12, per = 4 / tail;          // warps per tail strip: 4 or 2
    const int part = j % per;
    strip = 12 + j / per;
    first = (other * part) / per;
    cnt = (other * (part + 1)) / per - first;
  } else if (warp < own) {
    strip = warp; cnt = other;
  }
}

template <int NT>
__device__ __forceinline__ void stage1_strip(double (&acc)[MAXT][2], const double* __restrict__ F, long long rs, long long cs, int m1,
                                             int n1, int mt, int g, int t, int K1, const double* sm) {
  auto loadF = [&](int ks) -> double {
    const int r = 8 * mt + g, c = 4 * ks + t;
    return (r < m1 && c < n1) ? __ldg(F + r * rs + c * cs) : 0.0;
  };
  double a0 = loadF(0), a1 = K1 > 1 ? loadF(1) : 0.0;     // two k-steps of prefetch
  const double* xb = sm + t + PITCH * g;
  for (int ks = 0; ks < K1; ks++, xb += 4) {
    const double na = ks + 2 < K1 ? loadF(ks + 2) : 0.0;
#pragma unroll
    for (int nt = 0; nt < NT; nt++) dmma884(acc[nt][0], acc[nt][1], a0, xb[PITCH * 8 * nt]);
    a0 = a1; a1 = na;
  }
}
template <int MT>
__device__ __forceinline__ void stage2_strip(double (&z)[MAXT][2], const double* __restrict__ F, long long rs, long long cs, int m2,
                                             int n2, int nt, int g, int t, int K2, const double* sm) {
  auto loadF = [&](int ks) -> double {
    const int r = 8 * nt + g, c = 4 * ks + t;               // B[t][g] = F2op(b' = 8 nt + g, b = 4 ks + t)
    return (r < m2 && c < n2) ? __ldg(F + r * rs + c * cs) : 0.0;
  };
  double b0 = loadF(0), b1 = K2 > 1 ? loadF(1) : 0.0;
  const double* ya = sm + g + PITCH * t;
  for (int ks = 0; ks < K2; ks++, ya += 4 * PITCH) {
    const double nb = ks + 2 < K2 ? loadF(ks + 2) : 0.0;
#pragma unroll
    for (int mt = 0; mt < MT; mt++) dmma884(z[mt][0], z[mt][1], ya[8 * mt], b0);
    b0 = b1; b1 = nb;
  }
}
#define K2S_DISPATCH(N, CALL)                                                                                          \
  switch (N) {                                                                                                         \
    case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; case 4: CALL(4); break;                    \
    case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; case 8: CALL(8); break;                    \
    case 9: CALL(9); break; case 10: CALL(10); break; case 11: CALL(11); break; case 12: CALL(12); break;              \
    case 13: CALL(13); break; case 14: CALL(14); break; case 15: CALL(15); break; default: CALL(16); break;            \
  }

__global__ void __launch_bounds__(THREADS, 1) kron2_f64_small_kernel(const Params p) {
  extern __shared__ __align__(16) double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int MT1 = (p.m1 + 7) / 8, NT1 = (p.n2 + 7) / 8;     // stage 1 output strips: rows c', columns b
  const int MT2 = MT1, NT2 = (p.m2 + 7) / 8;                // stage 2 output strips: rows c', columns b'
  const int K1 = (p.n1 + 3) / 4, K2 = (p.n2 + 3) / 4;       // k-steps of 4
  const int n2p = NT1 * 8;

  for (long long s = blockIdx.x; s < p.slabs; s += gridDim.x) {
    // ---- X_s -> shared memory (pitch px, zero padding) ----
    __syncthreads();                                        // the previous slab's stage 2 has finished reading Y
    const double* Vs = p.V + s * (long long)p.n1 * p.n2;
    if (p.n1 % 2 == 0 && (reinterpret_cast<uintptr_t>(p.V) & 15u) == 0) {
      // every column of the slab is a run of n1 / 2 16-byte pieces: all of a thread's cp.async are in flight at once (one
      // memory latency for the whole slab instead of one per round of the register path below); the padding the DMMA
      // k-steps and strips read (rows n1 .. 4 K1 - 1, columns n2 .. n2p - 1) is zeroed by ordinary stores
      const int half = p.n1 >> 1, pieces = half * p.n2;
      const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
      for (int i = threadIdx.x; i < pieces; i += THREADS) {
        const int b = i / half, c = (i - b * half) * 2;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + (uint32_t)(c + PITCH * b) * 8u),
                     "l"(Vs + c + (long long)p.n1 * b)
                     : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      const int rpad = 4 * K1 - p.n1;                       // 0 .. 3 rows below the slab
      for (int i = threadIdx.x; i < rpad * p.n2; i += THREADS) sm[p.n1 + i % rpad + PITCH * (i / rpad)] = 0.0;
      const int cpad = n2p - p.n2, rows = 4 * K1;           // whole padded columns
      for (int i = threadIdx.x; i < cpad * rows; i += THREADS) sm[i % rows + PITCH * (p.n2 + i / rows)] = 0.0;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
      // eight independent loads in flight per thread (a rolled loop pays one L2 / HBM latency per element)
      const int total = PITCH * n2p;
      for (int i0 = threadIdx.x; i0 < total; i0 += 8 * THREADS) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int i = i0 + u * THREADS;
          const int c = i % PITCH, b = i / PITCH;
          v[u] = (i < total && c < p.n1 && b < p.n2) ? __ldg(Vs + c + (long long)p.n1 * b) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int i = i0 + u * THREADS;
          if (i < total) sm[i] = v[u];
        }
      }
    }
    __syncthreads();

    // ---- stage 1: warp w -> row strip mt = w of Y, all column strips ----
    double acc[MAXT][2];
#pragma unroll
    for (int j = 0; j < MAXT; j++) acc[j][0] = acc[j][1] = 0.0;
    int mtA, ntA0, ntAc;
    assign_strips(warp, MT1, NT1, mtA, ntA0, ntAc);
    const bool hasA = ntAc > 0;
    if (hasA) {
#define K2S_S1(NT) stage1_strip<NT>(acc, p.F1, p.f1_rs, p.f1_cs, p.m1, p.n1, mtA, g, t, K1, sm + PITCH * 8 * ntA0)
      K2S_DISPATCH(ntAc, K2S_S1)
#undef K2S_S1
    }
    __syncthreads();                                        // every warp is done reading X: Y may overwrite it
    if (hasA) {
#pragma unroll
      for (int nt = 0; nt < MAXT; nt++) {
        if (nt < ntAc) {
          double* y = sm + (8 * mtA + g) + PITCH * (8 * (ntA0 + nt) + 2 * t);
          y[0] = acc[nt][0];
          y[PITCH] = acc[nt][1];
        }
      }
    }
    __syncthreads();

    // ---- stage 2: warp w -> column strip nt = w of Z, all row strips ----
    double z[MAXT][2];
#pragma unroll
    for (int i = 0; i < MAXT; i++) z[i][0] = z[i][1] = 0.0;
    int ntA, mtB0, mtBc;
    assign_strips(warp, NT2, MT2, ntA, mtB0, mtBc);
    const bool hasNA = mtBc > 0;
    if (hasNA) {
#define K2S_S2(MT) stage2_strip<MT>(z, p.F2, p.f2_rs, p.f2_cs, p.m2, p.n2, ntA, g, t, K2, sm + 8 * mtB0)
      K2S_DISPATCH(mtBc, K2S_S2)
#undef K2S_S2
      // ---- epilogue: Z_s = alpha * Z + beta * Z_s (beta == 0: never read) ----
      double* Ws = p.W + s * (long long)p.m1 * p.m2;
#pragma unroll
      for (int mt = 0; mt < MAXT; mt++) {
        if (mt < mtBc) {
          const int r = 8 * (mtB0 + mt) + g;
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int c = 8 * ntA + 2 * t + e;
            if (r < p.m1 && c < p.m2) {
              double* w = Ws + r + (long long)p.m1 * c;
              double v = p.alpha * z[mt][e];
              if (p.beta != 0.0) v += p.beta * *w;
              *w = v;
            }
          }
        }
      }
    }
  }
}

}  // namespace k2s
}  // namespace scimlop
