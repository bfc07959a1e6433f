// kron2_f16.cuh -- TWO Kronecker mode products of small dense Float32 factors fused on-chip (config 3: kron(A, B, C)
// with 128 x 128 factors).  The reference applies one mode per pass with the tensor going back to memory in between
// (src/tensor.jl:543-610, 750-834; the LoopVectorization extension's fused nest covers one outer stage,
// ext/SciMLOperatorsLoopVectorizationExt.jl:17-23).  Here one CTA takes a slab X_s (n1 x n2, the two fastest modes of
// one (slower index, column) pair), and produces
//
//     Z_s = F1op * X_s * F2op^T            F1op: m1 x n1 (innermost factor),  F2op: m2 x n2 (next factor)
//
// without the n1 x n2 -> m1 x n2 intermediate ever leaving the SM: HBM sees X once and Z once.
//
// Arithmetic: both products run on tcgen05.mma kind::f16 with a two-term FP16 split of every operand
// (x = hi + lo, 11 + 11 significand bits, the same 22 bits 3xTF32 keeps) and FP32 accumulation in TMEM:
// lo*Fhi + hi*Flo + hi*Fhi.  kind::f16 issues at twice the kind::tf32 rate, which is what makes two chained
// products per slab fit under the slab's HBM time.  FP16 has 5 exponent bits, so every operand is first scaled by
// an exact power of two: the factors by their largest entry (once per CTA), each slab by its largest |x| (found by
// the splitter warps while the slab sits in their registers), the intermediate by a fixed 2^-8 that its bound
// guarantees.  The scales are undone exactly in the epilogue; an entry more than 2^21 below its slab's largest keeps
// fewer than 18 bits but its ABSOLUTE error stays below 2^-39 of that largest entry, so the norm-wise error is that of
// 3xTF32 (~1e-6, tests/test_gpu_kron2.py).
//
//   stage 1:  Y^T(b, c') = sum_c  X(c, b)  * F1op(c', c)     A = X^T  (TMEM, written by the splitters)   B = F1 image (smem, K-major)
//   stage 2:  Z^T(b', c') = sum_b F2op(b', b) * Y^T(b, c')    A = F2 image (TMEM, resident)               B = Y^T split (smem, MN-major)
//
// The A operand of a tcgen05.mma keeps its row index, so the streamed data has to be the B operand of the second
// product: drain-1 warps read the stage-1 accumulator, split it and write it to shared memory in the MN-major
// canonical layout.  TMEM (512 columns): [0,128) stage-1 accumulator, [128,256) stage-2 accumulator, [256,384) four
// k-blocks of X^T (16 columns hi + 16 columns lo each), [384,512) F2 image (64 hi + 64 lo).
//
// Warp roles (20 warps): 0 TMA producer, 1 MMA issuer, 4-11 splitters (group g converts k-blocks 2g, 2g+1 of every
// slab), 12-15 drain-1, 16-19 drain-2 (alpha / beta epilogue, bulk tensor stores).
#pragma once
#include <cuda_fp16.h>

#include "gemm_tf32.cuh"

namespace scimlop {
namespace kf16 {

using tf32::fence_async_smem;
using tf32::make_desc_k;
using tf32::make_desc_raw;
using tf32::tc_commit;
using tf32::tc_fence_after;
using tf32::tc_fence_before;
using tf32::tc_ld32;
using tf32::tc_st32;
using tf32::tc_wait_ld;
using tf32::tc_wait_st;

constexpr int THREADS = 640;
constexpr int XT_BYTES = 16384;         // one raw k-block of a slab: 128 rows (b) x 32 c, 128-byte rows, SWIZZLE_128B
constexpr int IMG_BYTES = 32768;        // one 128 x 128 FP16 image
constexpr int EPI_STAGE_BYTES = 4096;   // one 32 x 32 FP32 block; two per drain-2 warp
constexpr int NBAR = 2 * 4 + 2 + 5 + 8;
constexpr int SMEM_BYTES = 1024 + 2 * IMG_BYTES + 4 * XT_BYTES + 2 * IMG_BYTES + 8 * EPI_STAGE_BYTES + 8 * NBAR + 256;
constexpr int TM_ACC1 = 0, TM_ACC2 = 128, TM_XA = 256, TM_F2HI = 384, TM_F2LO = 448;
constexpr int K0 = 6;                   // a slab's largest |x| is scaled into [2^6, 2^7): FP16-normal down to 2^-20 of it
// then |Y'| = |sum_c X' F1'| < 128 * 2^7 * 2 = 2^15 with |F1'| < 2: the intermediate is inside FP16 range as it is
constexpr int CTRL_REGS = 40, SPLIT_REGS = 120;
constexpr int NC = 128;                 // UMMA N: a tcgen05.mma costs ~20 clk on top of N/2, so the widest N wins (measured)

// Build with -DSCIMLOP_KF16_TRACE (tools/kf16_trace.py) to record clock64() stamps of every role for the first slabs of
// CTA 0: Params::trace[64 * iteration + slot].  Compiled out of the shipped library.
#ifdef SCIMLOP_KF16_TRACE
#define KF16_STAMP(it, slot) do { if (p.trace && blockIdx.x == 0 && (it) < 64 && lane == 0) p.trace[64 * (it) + (slot)] = clock64(); } while (0)
#define KF16_STAMP_NS(it, slot) do { if (p.trace && blockIdx.x == 0 && (it) < 64 && lane == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.trace[64 * (it) + (slot)] = (long long)t_; } } while (0)
#else
#define KF16_STAMP(it, slot) do { } while (0)
#define KF16_STAMP_NS(it, slot) do { } while (0)
#endif

struct Params {
  long long* trace;
  int n1, n2, m1, m2;           // F1op: m1 x n1 (fastest mode), F2op: m2 x n2 (next mode); all <= 128
  long long slabs;              // (slower modes) x (columns)
  const float* F1;              // F1op(c', c) at c' * f1_rs + c * f1_cs
  long long f1_rs, f1_cs;
  const float* F2;
  long long f2_rs, f2_cs;
  float alpha, beta;
};

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// instruction descriptor: D = F32, A = B = F16, A K-major (TMEM), M = 128, N = 128
__device__ __forceinline__ uint32_t make_idesc_f16(bool b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                        // c_format = F32
  d |= (b_mn ? 1u : 0u) << 16;         // b_major (a_format = b_format = 0: F16)
  d |= (uint32_t)(NC >> 3) << 17;      // N >> 3
  d |= (uint32_t)(128 >> 4) << 24;     // M >> 4
  return d;
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 2^e as a float, e in [-126, 127]
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }
// unbiased exponent of the largest magnitude `bits` (|x| as an integer): 0 for zero / Inf / NaN, -126 for denormals
__device__ __forceinline__ int exp_of_max(uint32_t bits) {
  const int e = (int)(bits >> 23);
  if (bits == 0u || e == 255) return 0;
  return (e == 0 ? 1 : e) - 127;
}
// (a, b) -> packed FP16 hi pair and lo pair of the two-term split
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(THREADS, 1)
    kron2_f16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
  const uint32_t sF1hi = smem, sF1lo = smem + IMG_BYTES;          // stage-1 B operand: K-major, two 64-k panels of 16 KB
  const uint32_t sX = smem + 2 * IMG_BYTES;                       // 4 raw k-blocks
  const uint32_t sYhi = sX + 4 * XT_BYTES, sYlo = sYhi + IMG_BYTES;   // stage-2 B operand: MN-major, two 64-n halves
  const uint32_t sEpi = sYlo + IMG_BYTES;
  const uint32_t bars = sEpi + 8 * EPI_STAGE_BYTES;
  const uint32_t raw_full = bars, raw_free = raw_full + 32, a_full = raw_free + 32, a_free = a_full + 8;
  const uint32_t tmem1_full = a_free + 8, tmem1_empty = tmem1_full + 8, y_full = tmem1_empty + 8;
  const uint32_t tmem2_full = y_full + 8, tmem2_empty = tmem2_full + 8;
  const uint32_t epi_bar = tmem2_empty + 8;                       // [4 warps][2 buffers]
  const uint32_t misc = epi_bar + 64;                             // tmem slot, wmax[2][8], sexp[8], fmax[2], scratch[20]
  const uint32_t tmem_slot = misc, wmax = misc + 16, sexp = wmax + 64, fmax = sexp + 32, scratch = fmax + 16;
  uint32_t* misc_gen = reinterpret_cast<uint32_t*>(smem_gen + (misc - smem));
  volatile uint32_t* wmax_g = misc_gen + 4;
  volatile int* sexp_g = reinterpret_cast<volatile int*>(misc_gen + 20);
  uint32_t* fmax_g = misc_gen + 28;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; s++) {
      mbar_init(raw_full + 8 * s, 1);
      mbar_init(raw_free + 8 * s, 4);
    }
    mbar_init(a_full, 8);        // every splitter warp, once both its k-blocks are in TMEM
    mbar_init(a_free, 1);        // tcgen05.commit after stage 1
    mbar_init(tmem1_full, 1);
    mbar_init(tmem1_empty, 4);
    mbar_init(y_full, 4);
    mbar_init(tmem2_full, 1);
    mbar_init(tmem2_empty, 4);
    for (int q = 0; q < 8; q++) mbar_init(epi_bar + 8 * q, 1);
    fmax_g[0] = 0u;
    fmax_g[1] = 0u;
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmD);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // ---- factor images (once per CTA): largest entry -> exact power-of-two scale -> FP16 hi / lo ----
  {
    uint32_t m1b = 0u, m2b = 0u;
    for (int i = threadIdx.x; i < p.m1 * p.n1; i += THREADS) {
      const int r = i % p.m1, c = i / p.m1;
      m1b = max(m1b, __float_as_uint(__ldg(p.F1 + r * p.f1_rs + c * p.f1_cs)) & 0x7FFFFFFFu);
    }
    for (int i = threadIdx.x; i < p.m2 * p.n2; i += THREADS) {
      const int r = i % p.m2, c = i / p.m2;
      m2b = max(m2b, __float_as_uint(__ldg(p.F2 + r * p.f2_rs + c * p.f2_cs)) & 0x7FFFFFFFu);
    }
    m1b = __reduce_max_sync(0xffffffffu, m1b);
    m2b = __reduce_max_sync(0xffffffffu, m2b);
    if (lane == 0) {
      atomicMax(&fmax_g[0], m1b);
      atomicMax(&fmax_g[1], m2b);
    }
  }
  __syncthreads();
  const int e1 = exp_of_max(fmax_g[0]), e2 = exp_of_max(fmax_g[1]);
  {
    // F1 image: element (n = c', k = c) of the K-major SWIZZLE_128B canonical layout (rows of 64 k = 128 B)
    const float s1a = pow2f(-(e1 / 2)), s1b = pow2f(-(e1 - e1 / 2));   // 2^-e1 in two exact steps
    for (int i = threadIdx.x; i < 128 * 128; i += THREADS) {
      const int r = i & 127, c = i >> 7;
      float v = 0.f;
      if (r < p.m1 && c < p.n1) v = __ldg(p.F1 + r * p.f1_rs + c * p.f1_cs) * s1a * s1b;
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      const uint32_t off = (uint32_t)(c >> 6) * 16384u + (uint32_t)r * 128u + (uint32_t)((((c & 63) >> 3) ^ (r & 7)) << 4) + (uint32_t)(c & 7) * 2u;
      *reinterpret_cast<__half*>(smem_gen + off) = h;
      *reinterpret_cast<__half*>(smem_gen + IMG_BYTES + off) = l;
    }
    fence_async_smem();
  }
  if (warp >= 12 && warp < 16) {
    // F2 image into TMEM: lane = b', 32-bit column j = (F2op(b', 2j), F2op(b', 2j+1))
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const float s2a = pow2f(-(e2 / 2)), s2b = pow2f(-(e2 - e2 / 2));
    const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
    for (int t = 0; t < 2; t++) {
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const int b = 64 * t + 2 * j;
        float v0 = 0.f, v1 = 0.f;
        if (row < p.m2 && b < p.n2) v0 = __ldg(p.F2 + row * p.f2_rs + b * p.f2_cs) * s2a * s2b;
        if (row < p.m2 && b + 1 < p.n2) v1 = __ldg(p.F2 + row * p.f2_rs + (b + 1) * p.f2_cs) * s2a * s2b;
        split_pair(v0, v1, hi[j], lo[j]);
      }
      tc_st32(ta + TM_F2HI + 32 * t, hi);
      tc_st32(ta + TM_F2LO + 32 * t, lo);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int nkb1 = (p.n1 + 31) / 32;      // k-blocks of stage 1
  const int ks2 = (p.n2 + 15) / 16;       // k-steps of stage 2
  const int nblk = (p.m1 + 31) / 32;      // 32-column blocks of the output

  if (warp < 4) {
    setmaxnreg_dec<CTRL_REGS>();
    if (warp == 0) {
      // ===================== TMA producer =====================
      uint32_t ph = 0;
      long long it = 0;
      for (long long s = blockIdx.x; s < p.slabs; s += gridDim.x, ph ^= 1, it++) {
        for (int kk = 0; kk < nkb1; kk++) {
          mbar_wait_relaxed(raw_free + 8 * kk, ph ^ 1);
          KF16_STAMP(it, 0 + kk);
          if (elect_one()) {
            mbar_arrive_expect_tx(raw_full + 8 * kk, XT_BYTES);
            tma_load_3d(sX + kk * XT_BYTES, &tmX, raw_full + 8 * kk, 32 * kk, 0, (int)s);
          }
          __syncwarp();
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      const uint32_t idesc1 = make_idesc_f16(false), idesc2 = make_idesc_f16(true);
      const uint64_t f1hi = make_desc_k(sF1hi), f1lo = make_desc_k(sF1lo);
      const uint64_t yhi = make_desc_raw(sYhi, 16384, 1024, 2), ylo = make_desc_raw(sYlo, 16384, 1024, 2);
      uint32_t ph = 0;
      long long it = 0;
      // Every mbarrier wait costs this warp ~200 clk of shared-memory round trip (the pipe is ~75 % busy), so the waits
      // that are satisfied long before they are needed (accumulator / operand hand-backs) are placed right AFTER an issue
      // block, where the tensor pipe still has ~1900 clk of queued work; only y_full is a true dependency.
      mbar_wait(a_full, 0);
      for (long long s = blockIdx.x; s < p.slabs; s += gridDim.x, ph ^= 1, it++) {
        KF16_STAMP(it, 8);
        KF16_STAMP_NS(it, 50);
        tc_fence_after();
        if (elect_one()) {
          // stage 1: Y^T = X^T F1op^T, one uninterrupted issue block (a commit between k-blocks drains the pipe)
          const uint32_t d = tmem_base + TM_ACC1;
          for (int k0 = 0; k0 < p.n1; k0 += 16) {
            const uint32_t ahi = tmem_base + TM_XA + (k0 >> 4) * 8 + (k0 >> 5) * 16, alo = ahi + 16;
            const uint64_t koff = (uint64_t)(((k0 >> 6) * 16384 + (k0 & 63) * 2) >> 4);
            tc_mma_f16_ts(d, alo, f1hi + koff, idesc1, k0 == 0 ? 0u : 1u);
            tc_mma_f16_ts(d, ahi, f1lo + koff, idesc1, 1u);
            tc_mma_f16_ts(d, ahi, f1hi + koff, idesc1, 1u);
          }
          tc_commit(tmem1_full);
          tc_commit(a_free);
        }
        __syncwarp();
        KF16_STAMP(it, 14);
        mbar_wait(tmem2_empty, ph ^ 1);
        mbar_wait(y_full, ph);
        KF16_STAMP(it, 15);
        tc_fence_after();
        if (elect_one()) {
          // stage 2: Z^T = F2op Y^T
          const uint32_t d = tmem_base + TM_ACC2;
          for (int ks = 0; ks < ks2; ks++) {
            const uint32_t ahi = tmem_base + TM_F2HI + ks * 8, alo = tmem_base + TM_F2LO + ks * 8;
            const uint64_t koff = (uint64_t)((ks * 2048) >> 4);
            tc_mma_f16_ts(d, alo, yhi + koff, idesc2, ks == 0 ? 0u : 1u);
            tc_mma_f16_ts(d, ahi, ylo + koff, idesc2, 1u);
            tc_mma_f16_ts(d, ahi, yhi + koff, idesc2, 1u);
          }
          tc_commit(tmem2_full);
        }
        __syncwarp();
        KF16_STAMP(it, 16);
        if (s + gridDim.x < p.slabs) {
          mbar_wait(tmem1_empty, ph);        // drain-1 of THIS slab (it finished before stage 2 could start)
          mbar_wait(a_full, ph ^ 1);         // X^T of the next slab
        }
      }
    }
  } else if (warp < 12) {
    setmaxnreg_inc<SPLIT_REGS>();
    // ===================== splitters: raw slab (smem) -> slab scale -> FP16 hi / lo of X^T (TMEM) =====================
    const int grp = (warp - 4) >> 2, quarter = warp & 3, row = quarter * 32 + lane;
    uint32_t ph = 0;
    long long it = 0;
    for (long long s = blockIdx.x; s < p.slabs; s += gridDim.x, ph ^= 1, it++) {
      float x[64];
      uint32_t mx = 0u;
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int kk = 2 * grp + j;
        if (kk < nkb1) {
          mbar_wait(raw_full + 8 * kk, ph);
          if (quarter == 0) KF16_STAMP(it, 20 + kk);
          const uint8_t* g = smem_gen + (sX - smem) + kk * XT_BYTES + row * 128;
          uint32_t digest = 0;
#pragma unroll
          for (int cidx = 0; cidx < 8; cidx++) {
            const float4 v = *reinterpret_cast<const float4*>(g + ((cidx ^ (row & 7)) << 4));
            x[32 * j + 4 * cidx + 0] = v.x; x[32 * j + 4 * cidx + 1] = v.y;
            x[32 * j + 4 * cidx + 2] = v.z; x[32 * j + 4 * cidx + 3] = v.w;
            const uint32_t b0 = __float_as_uint(v.x), b1 = __float_as_uint(v.y), b2 = __float_as_uint(v.z), b3 = __float_as_uint(v.w);
            digest ^= b0 ^ b1 ^ b2 ^ b3;
            mx = max(max(mx, b0 & 0x7FFFFFFFu), max(b1 & 0x7FFFFFFFu, max(b2 & 0x7FFFFFFFu, b3 & 0x7FFFFFFFu)));
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_after_loads(raw_free + 8 * kk, scratch + 4 * warp, digest);
        } else {
#pragma unroll
          for (int q = 0; q < 32; q++) x[32 * j + q] = 0.f;
        }
      }
      // largest |x| of the slab over the 8 splitter warps
      mx = __reduce_max_sync(0xffffffffu, mx);
      const int slot = (int)(it & 1);
      if (lane == 0) wmax_g[8 * slot + (warp - 4)] = mx;
      if (quarter == 0) KF16_STAMP(it, 24 + grp);
      bar_sync_named(1, 256);
      if (quarter == 0) KF16_STAMP(it, 26 + grp);
      uint32_t all = 0u;
#pragma unroll
      for (int q = 0; q < 8; q++) all = max(all, wmax_g[8 * slot + q]);
      const int ex = exp_of_max(all);
      if (warp == 4 && lane == 0) sexp_g[it & 7] = ex;
      // x * 2^(K0 - ex) in two exact steps (the exponent K0 - ex spans -114 .. 140)
      const int sh = K0 - ex, sh1 = sh / 2;
      const float sa = pow2f(sh1), sb = pow2f(sh - sh1);
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int kk = 2 * grp + j;
        if (kk < nkb1) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int q = 0; q < 16; q++)
            split_pair(x[32 * j + 2 * q] * sa * sb, x[32 * j + 2 * q + 1] * sa * sb, hi[q], lo[q]);
          if (quarter == 0) KF16_STAMP(it, 28 + kk);
          if (j == 0) mbar_wait(a_free, ph ^ 1);   // stage 1 of the previous slab has read all of X^T
          if (quarter == 0) KF16_STAMP(it, 32 + kk);
          tc_fence_after();
          const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + TM_XA + kk * 32;
          tc_st16(ta, hi);
          tc_st16(ta + 16, lo);
          if (quarter == 0) KF16_STAMP(it, 36 + kk);
        }
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
      KF16_STAMP(it, 52 + 4 * grp + quarter);
    }
  } else if (warp < 16) {
    // ===================== drain-1: stage-1 accumulator -> FP16 hi / lo of Y^T, MN-major B operand of stage 2 =====================
    const int quarter = warp & 3, row = quarter * 32 + lane;       // row = b = k index of stage 2
    const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + TM_ACC1;
    uint8_t* yh = smem_gen + (sYhi - smem) + (row >> 3) * 1024 + (row & 7) * 128;
    uint8_t* yl = yh + IMG_BYTES;
    uint32_t ph = 0;
    long long it = 0;
    for (long long s = blockIdx.x; s < p.slabs; s += gridDim.x, ph ^= 1, it++) {
      mbar_wait(tmem1_full, ph);
      if (quarter == 0) KF16_STAMP(it, 40);
      tc_fence_after();
      // software-pipelined: the tcgen05.ld of block q + 1 is in flight while block q is split and stored
      uint32_t v[2][32];
      tc_ld32(tl, v[0]);
#pragma unroll
      for (int blk = 0; blk < 4; blk++) {
        tc_wait_ld();
        if (quarter == 0) KF16_STAMP(it, 60 + blk);
        if (blk + 1 < 4) tc_ld32(tl + (blk + 1) * 32, v[(blk + 1) & 1]);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 16; q++)
          split_pair(__uint_as_float(v[blk & 1][2 * q]), __uint_as_float(v[blk & 1][2 * q + 1]), hi[q], lo[q]);
        // columns n = 32 blk + 8 t .. + 7  ->  half n / 64, 16-byte chunk (n % 64) / 8, XOR (k & 7)
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const int n = 32 * blk + 8 * t;
          const uint32_t off = (uint32_t)(n >> 6) * 16384u + (uint32_t)((((n & 63) >> 3) ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(yh + off) = make_uint4(hi[4 * t], hi[4 * t + 1], hi[4 * t + 2], hi[4 * t + 3]);
          *reinterpret_cast<uint4*>(yl + off) = make_uint4(lo[4 * t], lo[4 * t + 1], lo[4 * t + 2], lo[4 * t + 3]);
        }
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(y_full);
        mbar_arrive(tmem1_empty);
      }
      if (quarter == 0) KF16_STAMP(it, 41);
    }
  } else {
    // ===================== drain-2: stage-2 accumulator -> scales undone, alpha / beta -> smem -> bulk tensor store =====================
    const int quarter = warp & 3;
    const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + TM_ACC2;
    const uint32_t stage_s = sEpi + quarter * 2 * EPI_STAGE_BYTES;
    uint8_t* stage_g = smem_gen + (stage_s - smem);
    const uint32_t ebar = epi_bar + 16 * quarter;
    uint32_t eph0 = 0, eph1 = 0;
    const bool readw = (p.beta != 0.f);
    const int rw = quarter * 32;                                   // first b' of this warp
    const bool active = rw < p.m2;
    uint32_t ph = 0;
    long long it = 0;
    for (long long s = blockIdx.x; s < p.slabs; s += gridDim.x, ph ^= 1, it++) {
      auto load_old = [&](int blk) {
        if (lane == 0) {
          const int buf = blk & 1;
          tma_store_wait_read();
          mbar_arrive_expect_tx(ebar + 8 * buf, EPI_STAGE_BYTES);
          tma_load_3d(stage_s + buf * EPI_STAGE_BYTES, &tmD, ebar + 8 * buf, blk * 32, rw, (int)s);
        }
      };
      if (readw && active) load_old(0);
      float fa = 0.f, fb = 0.f;
#pragma unroll
      for (int blk = 0; blk < 4; blk++) {
        if (blk == 0) {
          mbar_wait_relaxed(tmem2_full, ph);
          if (quarter == 0) KF16_STAMP(it, 44);
          tc_fence_after();
          // Z = Z'' * 2^(ex + e1 + e2 - K0), in two steps that each stay inside the FP32 exponent range
          int kexp = sexp_g[it & 7] + e1 + e2 - K0;
          kexp = max(-250, min(250, kexp));
          fa = pow2f(kexp / 2);
          fb = pow2f(kexp - kexp / 2) * p.alpha;
        }
        uint32_t v[32];
        const bool live = blk < nblk && active;
        if (live) {
          tc_ld32(tl + blk * 32, v);
          tc_wait_ld();
        }
        if (quarter == 0) KF16_STAMP(it, 4 + blk);
        if (blk == 3) {   // the whole accumulator is in registers / staged: hand it back before the last store
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem2_empty);
        }
        if (!live) continue;
        const int buf = blk & 1;
        uint8_t* sg = stage_g + buf * EPI_STAGE_BYTES;
        if (readw) {
          if (blk + 1 < nblk) load_old(blk + 1);
          if (buf == 0) { mbar_wait(ebar, eph0); eph0 ^= 1; } else { mbar_wait(ebar + 8, eph1); eph1 ^= 1; }
        } else {
          if (lane == 0) tma_store_wait_read1();
          __syncwarp();
        }
        if (quarter == 0 && blk == 2) KF16_STAMP(it, 17);
        // output is contiguous along c': smem [b' row][32 c'], 16-byte chunks XOR (row & 7) (SWIZZLE_128B)
#pragma unroll
        for (int q = 0; q < 8; q++) {
          float4* dst = reinterpret_cast<float4*>(sg + lane * 128 + ((q ^ (lane & 7)) << 4));
          float4 o = make_float4(__uint_as_float(v[4 * q]) * fa * fb, __uint_as_float(v[4 * q + 1]) * fa * fb,
                                 __uint_as_float(v[4 * q + 2]) * fa * fb, __uint_as_float(v[4 * q + 3]) * fa * fb);
          if (readw) {
            const float4 old = *dst;
            o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
          }
          *dst = o;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmD, stage_s + buf * EPI_STAGE_BYTES, blk * 32, rw, (int)s);
          tma_store_commit();
        }
        __syncwarp();
      }
      if (quarter == 0) KF16_STAMP(it, 45);
    }
    if (lane == 0) tma_store_wait_all();
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace kf16
}  // namespace scimlop
