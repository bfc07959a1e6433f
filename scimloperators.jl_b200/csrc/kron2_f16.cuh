// kron2_f16.cuh -- TWO Kronecker mode products of small dense Float32 factors fused on-chip (config 3: kron(A, B, C)
// with 128 x 128 factors).  The reference applies one mode per pass with the tensor going back to memory in between
// (src/tensor.jl:543-610, 750-834; the LoopVectorization extension's fused nest covers one outer stage,
// ext/SciMLOperatorsLoopVectorizationExt.jl:17-23).  Here one CTA takes a slab X_s (n1 x n2, the two fastest modes of
// one (slower index, column) pair), and produces
//
//     Z_s = F1op * X_s * F2op^T            F1op: m1 x n1 (innermost factor),  F2op: m2 x n2 (next factor)
//
// without the m1 x n2 intermediate ever leaving the SM: HBM sees X once and Z once.
//
// Arithmetic: both products run on tcgen05.mma kind::f16 with a two-term FP16 split of every operand
// (x = hi + lo, 11 + 11 significand bits, the same 22 bits 3xTF32 keeps) and FP32 accumulation in TMEM:
// lo*Fhi + hi*Flo + hi*Fhi.  kind::f16 issues at twice the kind::tf32 rate, which is what makes two chained
// products per slab fit under the slab's HBM time.  FP16 has 5 exponent bits, so every operand is first scaled by
// an exact power of two: the factors by their largest entry (once per CTA), each slab by its largest |x| (found by
// the splitter warps while the slab sits in their registers) into [2^6, 2^7), which also bounds the intermediate
// (|Y'| < 128 * 2^7 * 2 = 2^15) inside FP16 range.  The scales are undone exactly in the epilogue; an entry more than
// 2^20 below its slab's largest keeps fewer than 22 bits but its ABSOLUTE error stays below 2^-31 of that largest entry,
// so the norm-wise error is that of 3xTF32 (~1e-6, tests/test_gpu_kron2.py).
//
//   stage 1:  Y(c', b)  = sum_c F1op(c', c) * X(c, b)      A = F1 image (TMEM, resident)          B = X split (smem, K-major)
//   stage 2:  Z(c', b') = sum_b Y(c', b)   * F2op(b', b)   A = Y split (TMEM, written by drain-1)  B = F2 image (smem, K-major)
//
// The A operand of a tcgen05.mma keeps its row index and its K index runs along TMEM columns, so with F1 as A the
// stage-1 accumulator already has the layout stage 2 wants for ITS A operand: drain-1 warps read it (tcgen05.ld), split
// it and write the FP16 pair back IN PLACE (tcgen05.st), no shared memory involved.  The accumulators keep c' -- the
// index that is contiguous in memory -- on the TMEM lanes, so drain-2 stores its registers straight to global memory,
// 128 contiguous bytes per warp instruction, without a staging pass.  That leaves shared memory (the resource this
// kernel is bound by: ~0.8 wavefront per clock, profiles/r02_ncu_kron2_f16_*) with the TMA landing of X, one LDS + STS
// pass for the FP16 split of X, and the two B-operand streams of the tensor core.
//
// TMEM (512 columns): [0,128) and [128,256) two stage-1 accumulators / stage-2 A operands (stage 1 of slab i+1 runs
// while slab i is drained and multiplied again), [256,384) stage-2 accumulator, [384,512) F1 image (64 hi + 64 lo).
//
// Warp roles (20 warps): 0 TMA producer, 1 MMA issuer, 4-11 splitters (group g converts k-blocks 2g, 2g+1 of every
// slab), 12-15 drain-1, 16-19 drain-2 (alpha / beta epilogue).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gemm_tf32.cuh"

namespace scimlop {
namespace kf16 {

using tf32::fence_async_smem;
using tf32::make_desc_k;
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
constexpr int RS_MAX = 6;                // raw k-block ring: 6 stages (1.5 slabs) in the 3-arg kernel, 4 + the old-output ring in the 5-arg one
constexpr int NBAR = 2 * RS_MAX + 2 + 2 * 2 + 2 + 4;
constexpr int WT_BYTES = 16384;         // one 128 x 32 block of the old output (beta != 0)
constexpr int SMEM_BYTES = 1024 + 2 * IMG_BYTES + RS_MAX * XT_BYTES + 2 * IMG_BYTES + 8 * NBAR + 256;
static_assert(RS_MAX * XT_BYTES == 4 * XT_BYTES + 2 * WT_BYTES, "both kernels use the same shared-memory budget");
constexpr int TM_ACC1 = 0, TM_ACC2 = 256, TM_F1HI = 384, TM_F1LO = 448;
constexpr int K0 = 6;                   // a slab's largest |x| is scaled into [2^6, 2^7): FP16-normal down to 2^-20 of it
// setmaxnreg budget (96 x 640 registers at launch): 128 x 32 + 256 x 96 + 128 x 72 + 128 x 176 = 60416
constexpr int CTRL_REGS = 32, SPLIT_REGS = 96, DRAIN1_REGS = 72, DRAIN2_REGS = 176;

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
  float* D;                     // slab s at D + s * m1 * m2, column-major m1 x m2
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
// instruction descriptor: D = F32, A = B = F16, both K-major (A in TMEM), M = 128, N = 128.  A tcgen05.mma costs ~20 clk on
// top of its N / 2, so the widest N wins (measured: 79 clk at N = 128, 52 at N = 64).
__device__ __forceinline__ uint32_t make_idesc_f16(bool bf16) {
  uint32_t d = 0;
  d |= 1u << 4;                        // c_format = F32 (a_format = b_format = 0: F16, 1: BF16; a_major = b_major = 0: K)
  if (bf16) d |= (1u << 7) | (1u << 10);
  d |= (uint32_t)(128 >> 3) << 17;     // N >> 3
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
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// byte offset of element (row n, k) in a 128 x 128 FP16 K-major SWIZZLE_128B operand image: two 64-k panels of 16 KB,
// rows of 128 B, 16-byte chunks XOR (n & 7)
__device__ __forceinline__ uint32_t kmajor_off(int n, int k) {
  return (uint32_t)(k >> 6) * 16384u + (uint32_t)n * 128u + (uint32_t)((((k & 63) >> 3) ^ (n & 7)) << 4) + (uint32_t)(k & 7) * 2u;
}

// READW: the 5-arg form (beta != 0).  The slab-wide scale makes a splitter need ALL k-blocks of a slab before it can convert
// any, so the raw ring has to hold more than one slab for the TMA latency to stay out of the splitters' loop: 6 stages
// without the old-output ring, 4 with it (that kernel is bound by its 1.5x memory traffic anyway).
// BF16: the opt-in single-pass mode (SCIMLOP_PREC_BF16): operands rounded to BF16 (no split, no scaling: BF16 has the
// FP32 exponent range), one tcgen05.mma kind::f16 per k-step with BF16 formats, FP32 accumulation; rel. error ~4e-3.
template <bool READW, bool BF16 = false>
__global__ void __launch_bounds__(THREADS, 1)
    kron2_f16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
  const uint32_t sF2hi = smem, sF2lo = smem + IMG_BYTES;          // stage-2 B operand (resident)
  constexpr int RS = READW ? 4 : RS_MAX;
  const uint32_t sX = smem + 2 * IMG_BYTES;                       // RS raw k-blocks
  const uint32_t sW = sX + 4 * XT_BYTES;                          // READW: two 128 x 32 blocks of the old output in flight
  const uint32_t sXhi = sX + RS_MAX * XT_BYTES, sXlo = sXhi + IMG_BYTES;   // stage-1 B operand: the split slab
  const uint32_t bars = sXlo + IMG_BYTES;
  const uint32_t raw_full = bars, raw_free = raw_full + 8 * RS_MAX, x_full = raw_free + 8 * RS_MAX, x_free = x_full + 8;
  const uint32_t acc1_full = x_free + 8, y_full = acc1_full + 16;  // [2 accumulators]
  const uint32_t acc2_full = y_full + 16, acc2_free = acc2_full + 8;
  const uint32_t w_full = acc2_free + 8, w_free = w_full + 16;
  const uint32_t misc = w_free + 16;                            // tmem slot, wmax[2][8], sexp[8], fmax[2], scratch[20]
  const uint32_t tmem_slot = misc, scratch = misc + 128;
  uint32_t* misc_gen = reinterpret_cast<uint32_t*>(smem_gen + (misc - smem));
  volatile uint32_t* wmax_g = misc_gen + 4;
  volatile int* sexp_g = reinterpret_cast<volatile int*>(misc_gen + 20);
  uint32_t* fmax_g = misc_gen + 28;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < RS; s++) {
      mbar_init(raw_full + 8 * s, 1);
      mbar_init(raw_free + 8 * s, 4);
    }
    mbar_init(x_full, 8);        // every splitter warp, once both its k-blocks are in shared memory
    mbar_init(x_free, 1);        // tcgen05.commit after stage 1
    for (int a = 0; a < 2; a++) {
      mbar_init(acc1_full + 8 * a, 1);
      mbar_init(y_full + 8 * a, 4);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_free, 4);
    for (int a = 0; a < 2; a++) {
      mbar_init(w_full + 8 * a, 1);
      mbar_init(w_free + 8 * a, 4);
    }
    fmax_g[0] = 0u;
    fmax_g[1] = 0u;
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
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
  if (!BF16) {
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
  const int e1 = BF16 ? 0 : exp_of_max(fmax_g[0]), e2 = BF16 ? 0 : exp_of_max(fmax_g[1]);
  {
    // F2 image: B operand of stage 2, element (n = b', k = b) = F2op(b', b) * 2^-e2
    const float sa = pow2f(-(e2 / 2)), sb = pow2f(-(e2 - e2 / 2));   // 2^-e2 in two exact steps
    for (int i = threadIdx.x; i < 128 * 128; i += THREADS) {
      const int r = i & 127, c = i >> 7;
      float v = 0.f;
      if (r < p.m2 && c < p.n2) v = __ldg(p.F2 + r * p.f2_rs + c * p.f2_cs) * sa * sb;
      const uint32_t off = kmajor_off(r, c);
      if (BF16) {
        *reinterpret_cast<__nv_bfloat16*>(smem_gen + off) = __float2bfloat16_rn(v);
      } else {
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        *reinterpret_cast<__half*>(smem_gen + off) = h;
        *reinterpret_cast<__half*>(smem_gen + IMG_BYTES + off) = l;
      }
    }
    fence_async_smem();
  }
  if (warp >= 12 && warp < 16) {
    // F1 image into TMEM: A operand of stage 1, lane = c', 32-bit column j = (F1op(c', 2j), F1op(c', 2j+1)) * 2^-e1
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const float sa = pow2f(-(e1 / 2)), sb = pow2f(-(e1 - e1 / 2));
    const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
    for (int t = 0; t < 2; t++) {
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const int c = 64 * t + 2 * j;
        float v0 = 0.f, v1 = 0.f;
        if (row < p.m1 && c < p.n1) v0 = __ldg(p.F1 + row * p.f1_rs + c * p.f1_cs) * sa * sb;
        if (row < p.m1 && c + 1 < p.n1) v1 = __ldg(p.F1 + row * p.f1_rs + (c + 1) * p.f1_cs) * sa * sb;
        if (BF16) { hi[j] = pack_bf16(v0, v1); lo[j] = 0u; }
        else split_pair(v0, v1, hi[j], lo[j]);
      }
      tc_st32(ta + TM_F1HI + 32 * t, hi);
      tc_st32(ta + TM_F1LO + 32 * t, lo);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int nkb1 = (p.n1 + 31) / 32;      // raw k-blocks of a slab
  const int nblk2 = (p.m2 + 31) / 32;     // 32-column blocks of the output
  const long long nmine = (p.slabs - blockIdx.x + gridDim.x - 1) / gridDim.x;   // slabs of this CTA: blockIdx.x + i * gridDim.x

  if (warp < 4) {
    setmaxnreg_dec<CTRL_REGS>();
    if (warp == 0) {
      // ===================== TMA producer =====================
      long long n = 0;      // k-blocks issued: stage n % RS, use n / RS
      for (long long i = 0; i < nmine; i++) {
        const long long s = blockIdx.x + i * gridDim.x;
        for (int kk = 0; kk < nkb1; kk++, n++) {
          const uint32_t stage = (uint32_t)(n % RS), use = (uint32_t)(n / RS);
          mbar_wait_relaxed(raw_free + 8 * stage, (use & 1) ^ 1);
          KF16_STAMP(i, 0 + kk);
          if (elect_one()) {
            mbar_arrive_expect_tx(raw_full + 8 * stage, XT_BYTES);
            tma_load_3d(sX + stage * XT_BYTES, &tmX, raw_full + 8 * stage, 32 * kk, 0, (int)s);
          }
          __syncwarp();
        }
      }
    } else if (warp == 2 && READW) {
      // ===================== TMA producer of the old output (5-arg form): 128 x 32 blocks, two in flight =====================
      long long n = 0;
      for (long long i = 0; i < nmine; i++) {
        const long long s = blockIdx.x + i * gridDim.x;
        for (int blk = 0; blk < nblk2; blk++, n++) {
          const uint32_t buf = (uint32_t)(n & 1);
          mbar_wait_relaxed(w_free + 8 * buf, (uint32_t)((n >> 1) & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(w_full + 8 * buf, WT_BYTES);
            tma_load_3d(sW + buf * WT_BYTES, &tmW, w_full + 8 * buf, 0, 32 * blk, (int)s);
          }
          __syncwarp();
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      // Issue order: S1(0), then per slab i: S1(i+1), S2(i).  Stage 1 of the next slab fills the other accumulator while
      // drain-1 turns slab i's accumulator into the A operand of its stage 2, so the tensor pipe always has the next
      // block queued and the ~200 clk round trip of every mbarrier wait (the shared-memory pipe is ~80 % busy) hides
      // behind ~1900 clk of queued MMAs.
      const uint32_t idesc = make_idesc_f16(BF16);
      const uint64_t xhi = make_desc_k(sXhi), xlo = make_desc_k(sXlo), f2hi = make_desc_k(sF2hi), f2lo = make_desc_k(sF2lo);
      auto stage1 = [&](long long i) {
        mbar_wait(x_full, (uint32_t)(i & 1));
        KF16_STAMP(i, 8);
        if (i == 0) KF16_STAMP_NS(i, 50);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + TM_ACC1 + (uint32_t)(i & 1) * 128;
          for (int k0 = 0; k0 < p.n1; k0 += 16) {
            const uint32_t ahi = tmem_base + TM_F1HI + (k0 >> 1), alo = tmem_base + TM_F1LO + (k0 >> 1);
            const uint64_t koff = (uint64_t)(((k0 >> 6) * 16384 + (k0 & 63) * 2) >> 4);
            if (BF16) {
              tc_mma_f16_ts(d, ahi, xhi + koff, idesc, k0 == 0 ? 0u : 1u);
            } else {
              tc_mma_f16_ts(d, alo, xhi + koff, idesc, k0 == 0 ? 0u : 1u);
              tc_mma_f16_ts(d, ahi, xlo + koff, idesc, 1u);
              tc_mma_f16_ts(d, ahi, xhi + koff, idesc, 1u);
            }
          }
          tc_commit(acc1_full + 8 * (uint32_t)(i & 1));
          tc_commit(x_free);
        }
        __syncwarp();
        KF16_STAMP(i, 9);
      };
      stage1(0);
      for (long long i = 0; i < nmine; i++) {
        if (i + 1 < nmine) stage1(i + 1);
        mbar_wait(acc2_free, (uint32_t)(i & 1) ^ 1);
        mbar_wait(y_full + 8 * (uint32_t)(i & 1), (uint32_t)((i >> 1) & 1));
        KF16_STAMP(i, 15);
        KF16_STAMP_NS(i, 51);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t y = tmem_base + TM_ACC1 + (uint32_t)(i & 1) * 128;
          for (int k0 = 0; k0 < p.n2; k0 += 16) {
            const uint32_t ahi = y + (k0 >> 5) * 32 + ((k0 >> 4) & 1) * 8, alo = ahi + 16;
            const uint64_t koff = (uint64_t)(((k0 >> 6) * 16384 + (k0 & 63) * 2) >> 4);
            if (BF16) {
              tc_mma_f16_ts(tmem_base + TM_ACC2, ahi, f2hi + koff, idesc, k0 == 0 ? 0u : 1u);
            } else {
              tc_mma_f16_ts(tmem_base + TM_ACC2, alo, f2hi + koff, idesc, k0 == 0 ? 0u : 1u);
              tc_mma_f16_ts(tmem_base + TM_ACC2, ahi, f2lo + koff, idesc, 1u);
              tc_mma_f16_ts(tmem_base + TM_ACC2, ahi, f2hi + koff, idesc, 1u);
            }
          }
          tc_commit(acc2_full);
        }
        __syncwarp();
        KF16_STAMP(i, 16);
      }
    }
  } else if (warp < 12) {
    setmaxnreg_inc<SPLIT_REGS>();
    // ===================== splitters: raw slab (smem) -> slab scale -> FP16 hi / lo of X, stage-1 B operand (smem) =====================
    const int grp = (warp - 4) >> 2, quarter = warp & 3, row = quarter * 32 + lane;     // row = b
    uint8_t* xh = smem_gen + (sXhi - smem) + grp * 16384 + row * 128;                   // k-blocks 2g, 2g+1 = 64-k panel g
    uint8_t* xl = xh + IMG_BYTES;
    uint32_t ph = 0;
    for (long long it = 0; it < nmine; it++, ph ^= 1) {
      float x[64];
      uint32_t mx = 0u;
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int kk = 2 * grp + j;
        if (kk < nkb1) {
          // A stage is served by both groups in turn; the named barrier below keeps them within one slab of each other,
          // so a parity wait can never be two phases behind.
          const long long n = it * nkb1 + kk;
          const uint32_t stage = (uint32_t)(n % RS), use = (uint32_t)(n / RS);
          mbar_wait_relaxed(raw_full + 8 * stage, use & 1);
          if (quarter == 0) KF16_STAMP(it, 20 + kk);
          const uint8_t* g = smem_gen + (sX - smem) + stage * XT_BYTES + row * 128;
#pragma unroll
          for (int cidx = 0; cidx < 8; cidx++) {
            const float4 v = *reinterpret_cast<const float4*>(g + ((cidx ^ (row & 7)) << 4));
            x[32 * j + 4 * cidx + 0] = v.x; x[32 * j + 4 * cidx + 1] = v.y;
            x[32 * j + 4 * cidx + 2] = v.z; x[32 * j + 4 * cidx + 3] = v.w;
            const uint32_t b0 = __float_as_uint(v.x), b1 = __float_as_uint(v.y), b2 = __float_as_uint(v.z), b3 = __float_as_uint(v.w);
            mx = max(max(mx, b0 & 0x7FFFFFFFu), max(b1 & 0x7FFFFFFFu, max(b2 & 0x7FFFFFFFu, b3 & 0x7FFFFFFFu)));
          }
          __syncwarp();
          // the running maximum depends on every loaded register: it is the digest that holds the arrive back
          if (lane == 0) mbar_arrive_after_loads(raw_free + 8 * stage, scratch + 4 * warp, mx);
        } else {
#pragma unroll
          for (int q = 0; q < 32; q++) x[32 * j + q] = 0.f;
        }
      }
      uint32_t hi[32], lo[32];
      if (BF16) {
        // no slab scale to agree on, but the two groups still meet here once per slab: ring stages are served by both
        // groups in turn, and a parity wait is only safe while they stay within one slab of each other
        bar_sync_named(1, 256);
#pragma unroll
        for (int q = 0; q < 32; q++) { hi[q] = pack_bf16(x[2 * q], x[2 * q + 1]); lo[q] = 0u; }
      } else {
        // largest |x| of the slab over the 8 splitter warps
        mx = __reduce_max_sync(0xffffffffu, mx);
        const int slot = (int)(it & 1);
        if (lane == 0) wmax_g[8 * slot + (warp - 4)] = mx;
        bar_sync_named(1, 256);
        uint32_t all = 0u;
#pragma unroll
        for (int q = 0; q < 8; q++) all = max(all, wmax_g[8 * slot + q]);
        const int ex = exp_of_max(all);
        if (warp == 4 && lane == 0) sexp_g[it & 7] = ex;
        // x * 2^(K0 - ex): the exponent spans -122 .. 132, so two exact steps (the first is 1 whenever one step can do it)
        const int sh = K0 - ex, sh1 = (sh >= -126 && sh <= 127) ? 0 : sh / 2;
        const float sa = pow2f(sh1), sb = pow2f(sh - sh1);
        // split before the wait (the FP32 values die here: 64 registers of floats become 32 + 32 of packed halves), so that
        // the section between "stage 1 of the previous slab has read the split slab" and "the next one is in place" -- the
        // one dependency of the MMA stream on this warp -- is 16 stores long
        if (sh1 == 0) {
#pragma unroll
          for (int q = 0; q < 32; q++) split_pair(x[2 * q] * sb, x[2 * q + 1] * sb, hi[q], lo[q]);
        } else {
#pragma unroll
          for (int q = 0; q < 32; q++) split_pair(x[2 * q] * sa * sb, x[2 * q + 1] * sa * sb, hi[q], lo[q]);
        }
      }
      if (quarter == 0) KF16_STAMP(it, 28 + grp);
      mbar_wait_relaxed(x_free, ph ^ 1);
      if (quarter == 0) KF16_STAMP(it, 32 + grp);
      // row b of panel g: 8 chunks of 8 k, chunk j at (j ^ (b & 7)) * 16
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t off = (uint32_t)((j ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(xh + off) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
        if (!BF16) *reinterpret_cast<uint4*>(xl + off) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_full);
      KF16_STAMP(it, 52 + 4 * grp + quarter);
    }
  } else if (warp < 16) {
    setmaxnreg_dec<DRAIN1_REGS>();
    // ===================== drain-1: stage-1 accumulator -> FP16 hi / lo of Y, written back in place as the stage-2 A operand =====================
    const int quarter = warp & 3;
    const uint32_t tl0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + TM_ACC1;
    for (long long it = 0; it < nmine; it++) {
      const uint32_t tl = tl0 + (uint32_t)(it & 1) * 128;
      mbar_wait_relaxed(acc1_full + 8 * (uint32_t)(it & 1), (uint32_t)((it >> 1) & 1));
      if (quarter == 0) KF16_STAMP(it, 40);
      tc_fence_after();
      // columns b = 32 blk .. 32 blk + 31 (FP32)  ->  16 columns of hi pairs, 16 columns of lo pairs, same 32 columns
#pragma unroll
      for (int blk = 0; blk < 4; blk++) {
        uint32_t v[32];
        tc_ld32(tl + blk * 32, v);
        tc_wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
          if (BF16) hi[q] = pack_bf16(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
          else split_pair(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]), hi[q], lo[q]);
        }
        tc_st16(tl + blk * 32, hi);
        if (!BF16) tc_st16(tl + blk * 32 + 16, lo);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(y_full + 8 * (uint32_t)(it & 1));
      if (quarter == 0) KF16_STAMP(it, 41);
    }
  } else {
    // ===================== drain-2: stage-2 accumulator -> scales undone, alpha / beta -> global memory =====================
    // lane = c' is the contiguous index of the output: one warp store instruction writes 128 contiguous bytes.  The whole
    // accumulator row (128 columns) is pulled into registers first and handed back, so that stage 2 of the next slab never
    // waits for this slab's stores (they are paced by the memory system, ~4000 clk per slab).
    setmaxnreg_inc<DRAIN2_REGS>();
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + TM_ACC2;
    const bool row_ok = row < p.m1;
    constexpr bool readw = READW;
    const long long slab_elems = (long long)p.m1 * p.m2;
    uint32_t ph = 0;
    long long wn = 0;      // blocks of the old output consumed (beta != 0)
    for (long long it = 0; it < nmine; it++, ph ^= 1) {
      const long long s = blockIdx.x + it * gridDim.x;
      float* out = p.D + s * slab_elems + row;
      mbar_wait_relaxed(acc2_full, ph);
      if (quarter == 0) KF16_STAMP(it, 44);
      tc_fence_after();
      // Z = alpha * Z'' * 2^kexp, kexp = ex + e1 + e2 - K0.  One multiplier when 2^kexp * alpha cannot over- / underflow on
      // its own (|kexp| <= 100: whatever it loses the product loses too, since 1 <= |Z''| < 2^23 for the entries that
      // matter); two exact steps otherwise.
      int kexp = BF16 ? 0 : sexp_g[it & 7] + e1 + e2 - K0;
      kexp = max(-250, min(250, kexp));
      const bool one = kexp >= -100 && kexp <= 100;
      const float fa = one ? 1.f : pow2f(kexp / 2), fb = (one ? pow2f(kexp) : pow2f(kexp - kexp / 2)) * p.alpha;
      uint32_t v[4][32];
#pragma unroll
      for (int blk = 0; blk < 4; blk++) tc_ld32(tl + blk * 32, v[blk]);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_free);
      if (quarter == 0) KF16_STAMP(it, 46);
      {
        float* o = out;
        const long long st = p.m1;
#pragma unroll
        for (int blk = 0; blk < 4; blk++) {
          if (32 * blk >= p.m2) break;
          float old[32];
          if (readw) {
            // the old 128 x 32 block was prefetched into shared memory (lane = c': conflict-free), released once read
            const uint32_t buf = (uint32_t)(wn & 1);
            mbar_wait(w_full + 8 * buf, (uint32_t)((wn >> 1) & 1));
            const float* wsrc = reinterpret_cast<const float*>(smem_gen + (sW - smem) + buf * WT_BYTES) + row;
            uint32_t digest = 0;
#pragma unroll
            for (int q = 0; q < 32; q++) {
              old[q] = wsrc[q * 128];
              digest ^= __float_as_uint(old[q]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_after_loads(w_free + 8 * buf, scratch + 4 * warp, digest);
            wn++;
          }
          if (!row_ok) continue;
          if (32 * blk + 32 <= p.m2 && one) {
            // full block, straight-line: FMUL / FFMA + STG + pointer step per column
#pragma unroll
            for (int q = 0; q < 32; q++, o += st) {
              float r = __uint_as_float(v[blk][q]) * fb;
              if (readw) r += p.beta * old[q];
              *o = r;
            }
          } else {
            // ragged last block, or a scale outside the single-multiplier range
#pragma unroll
            for (int q = 0; q < 32; q++, o += st) {
              if (32 * blk + q < p.m2) {
                float r = __uint_as_float(v[blk][q]) * fa * fb;
                if (readw) r += p.beta * old[q];
                *o = r;
              }
            }
          }
        }
      }
      if (quarter == 0) KF16_STAMP(it, 45);
    }
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
