// gemm_tf32.cuh -- Float32 GEMM on the 5th-generation tensor cores: tcgen05.mma (kind::tf32) with TMEM
// accumulators, TMA-staged operands, 3xTF32 split for FP32-grade accuracy.  Two modes of one kernel: "streaming
// operand x resident factor" (described first; the Kronecker mode products of config 3) and, with STREAM_F, the
// general GEMM in which the second operand streams too (see the comment above the kernel).
//
//   D(r, c) = alpha * sum_k X(r, k) * Fop(c, k)  +  beta * D(r, c)            r < rows, c < m, k < n
//
// X is the big streamed operand (a mode-unfolding of V, src/tensor.jl:558,780), Fop = op(F) is one small
// Kronecker factor / MatrixOperator (m x n, both <= 128) that stays resident in shared memory, split once
// per CTA into F_hi + F_lo.  Each X tile is split on the fly (x = hi + lo, both exactly representable in
// TF32) and the product is accumulated as  lo*Fhi + hi*Flo + hi*Fhi  in FP32 inside TMEM (rel. error ~1e-6,
// inside the 1e-5 budget of the north_star; the dropped lo*lo term is O(2^-22)).
//
// Two geometries, both pure tensor-map / descriptor choices (no data is ever transposed in memory):
//   XK  (innermost Kronecker mode, MatrixOperator mul!):  X(r,k) at k + ldx*r      -> K-major A operand
//   XMN (any other Kronecker mode, batched):              X(r,k) at r + ldx*k (+b) -> MN-major A operand
// and the factor as B operand: MN-major for a plain factor (F(c,k) at c + ldf*k), K-major for a transposed one.
//
// The split X operand is handed to the tensor core through TENSOR MEMORY (tcgen05.mma with the A operand in
// TMEM, written by tcgen05.st), not through shared memory: shared memory then only holds the resident factor
// (128 KB), a 4-deep ring of raw 16 KB X tiles whose slots are released as soon as the splitters have read
// them, and 16 KB of staging for the bulk tensor stores of the epilogue.  (A first version kept hi/lo in
// shared memory: 3 stages only, latency-bound at 45 % of HBM; profiles/r01_ncu_tf32_v1_summary.txt.)
//
// TMEM map (512 columns): [0,128) accumulator 0, [128,256) accumulator 1, [256,512) four A stages of
// 64 columns (32 k of X_hi, 32 k of X_lo; lane = row of the tile, column = k).
//
// Warp roles (20 warps): 0 TMA producer, 1 MMA issuer (+TMEM owner), two splitter groups of 4 warps that
// alternate over k-blocks (their LDS -> ALU -> tcgen05.st -> barrier chains overlap), two epilogue groups of 4
// warps (group e drains accumulator e), so no role is a serial latency chain on the tile path.
#pragma once
#include "ptx.cuh"

namespace scimlop {
namespace tf32 {

constexpr int TM = 128;                 // rows of X per tile = UMMA M = TMEM lanes
constexpr int BK = 32;                  // k per stage (128 bytes)
constexpr int MAXF = 128;               // factor extents supported (m, n <= 128)
constexpr int RSTAGES = 4;              // raw X tiles in flight (shared memory ring)
constexpr int ASTAGES = 4;              // split A tiles in flight (tensor memory ring)
constexpr int XT_BYTES = TM * BK * 4;   // 16 KB: one raw X tile
constexpr int F_BYTES = MAXF * MAXF * 4;  // 64 KB each for F_hi and F_lo
constexpr int NBAR = 2 * RSTAGES + 2 * ASTAGES + 4 + 1 + 8;
constexpr int EPI_STAGE_BYTES = 4096;   // one 32 x 32 FP32 block; two per epilogue warp (double-buffered)
constexpr int SMEM_BYTES = 1024 + 2 * F_BYTES + RSTAGES * XT_BYTES + 8 * EPI_STAGE_BYTES + 8 * NBAR + 16 + 64;
// The tensor core adds into its FP32 accumulator with round-toward-zero: on same-sign data the error of a long
// accumulation chain grows LINEARLY (measured: 2e-8 * K relative, 1e-5 at K = 512, 1.6e-4 at K = 8192 -- tools/
// tf32_accuracy.py).  So a TMEM accumulator only ever holds CHUNK_KB k-blocks (128 k = 48 MMAs); the epilogue warps
// drain every chunk into FP32 registers with round-to-nearest adds while the MMAs of the next chunk fill the other
// accumulator.  Error after that: ~1e-6 at any K.
constexpr int CHUNK_KB = 4;             // (streamed mode only: the resident-factor mode has K <= 128 = one chunk)
constexpr int EPI_REGS = 216, CTRL_REGS = 40;   // setmaxnreg: 128 running sums per epilogue thread (40 + 2*128 + 216 = 512)
constexpr int NSG = 2;                  // splitter groups (4 warps each) alternating over k-blocks
constexpr int NEG = 1;                  // epilogue groups (4 warps each)
// A ring slot must always be served by the SAME splitter group: a group that never waited on phase j of a
// barrier can pass a parity wait for phase j+1 spuriously while the barrier is still in phase j (parity waits
// only distinguish "one ahead" from "one behind").  RSTAGES = 5 with two groups faulted exactly like that.
static_assert(RSTAGES % NSG == 0 && ASTAGES % NSG == 0, "ring depths must be multiples of the splitter group count");
constexpr int THREADS = 32 * (4 + 4 * NSG + 4 * NEG);   // warps 0 producer, 1 MMA, 2-3 idle, then the groups
constexpr int TMEM_COLS = 512;          // 2 x 128 accumulator columns + 4 x 64 A-operand columns
constexpr int TMEM_A0 = 256;

struct Params {
  long long rows;        // extent of r per batch
  int m, n;              // factor extents (output columns, contraction length)
  int batch;
  int tiles_r;           // ceil(rows / 128)
  int tiles_c;           // STREAM_F: ceil(m / 128) column tiles; 1 otherwise
  int c_fast;            // tile order: column tiles fastest (concurrent CTAs share an X row tile through L2; right when
                         // the pre-split operand is L2-resident), else row tiles fastest (they share an F panel)
  long long total_tiles; // tiles_r * tiles_c * batch
  float* D;
  long long d_rs, d_cs, d_bs;  // D(r, c, b) at r*d_rs + c*d_cs + b*d_bs
  float alpha, beta;
  int vec_ok;            // d_cs == 1 and 16-byte alignment holds: float4 stores along c
  int x_batched;
  int passes;            // 3: lo*Fhi + hi*Flo + hi*Fhi (FP32-grade);  1: hi*Fhi only (plain TF32, opt-in)
  // streamed mode: up to 3 row scalings of the product (left Diagonal / BatchedDiagonal factors of a tree term):
  // value(m, n) *= scale[q][m + (ld ? n * ld : 0)], with (m, n) = (r, c) or, when rows_are_n, (c, r)
  int nscale, rows_are_n;
  const float* scale_ptr[3];
  long long scale_ld[3];
  int tma_store;         // D is TMA-addressable: epilogue goes TMEM -> smem (-> + beta * old block, TMA-loaded) -> bulk store
};

// ---- tcgen05 wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// A operand from tensor memory (lane = row, column = k), B from a shared-memory descriptor
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 64-bit shared-memory matrix descriptor (SM100 format, version 1), SWIZZLE_128B.
//   K-major : rows of 128 B, 8-row groups SBO bytes apart; LBO unused (1).
//   MN-major: atoms of (32 elements along MN = 128 B) x (8 k); LBO = bytes between atoms along MN,
//             SBO = bytes between atoms along K.
//
// 32-bit MN-major operands have exactly one legal layout: SWIZZLE_128B_BASE32B (layout type 1): rows of
// 128 B (32 elements along MN) per k, 32-byte chunks XOR (k & 3), K atoms of 4 rows (SBO = 512 B when the k rows
// are contiguous), LBO = bytes between successive 32-element MN groups.  TMA writes it with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  (Plain SWIZZLE_128B silently yields zeros: tools/tc_probe.cu.)
__device__ __forceinline__ uint64_t make_desc_raw(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;       // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ uint64_t make_desc_k(uint32_t addr) { return make_desc_raw(addr, 16, 1024, 2); }
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t addr, uint32_t lbo_bytes) { return make_desc_raw(addr, lbo_bytes, 512, 1); }
// instruction descriptor: D = F32, A = B = TF32, M = 128, N = n_mma
__device__ __forceinline__ uint32_t make_idesc(bool a_mn, bool b_mn, int n_mma) {
  uint32_t d = 0;
  d |= 1u << 4;                        // c_format  = F32
  d |= 2u << 7;                        // a_format  = TF32
  d |= 2u << 10;                       // b_format  = TF32
  d |= (a_mn ? 1u : 0u) << 15;         // a_major   (0 = K, 1 = MN)
  d |= (b_mn ? 1u : 0u) << 16;         // b_major
  d |= (uint32_t)(n_mma >> 3) << 17;   // N >> 3
  d |= (uint32_t)(TM >> 4) << 24;      // M >> 4
  return d;
}

__device__ __forceinline__ void split4(float4 x, float4& hi, float4& lo) {
  auto h = [](float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); };
  hi = make_float4(h(x.x), h(x.y), h(x.z), h(x.w));
  lo = make_float4(h(x.x - hi.x), h(x.y - hi.y), h(x.z - hi.z), h(x.w - hi.w));
}

// X_MN: X is MN-major (r contiguous); else K-major (k contiguous).  F_K: factor is K-major (transposed factor).
//
// STREAM_F = true is the general GEMM (any contraction length, any number of output columns): the second operand
// is no longer a small resident factor but streams through the SAME ring as X, as 128-column x 32-k tiles of its
// PRE-SPLIT halves F_hi / F_lo (tmF / tmF2; split once per call by split_hilo_kernel, so that the B operand needs no
// in-kernel shared->shared pass).  A ring stage is then X | F_hi | F_lo (3 x 16 KB) and is released by the four
// splitter warps AND by the tcgen05.commit of the MMAs that read it.
template <bool X_MN, bool F_K, bool STREAM_F = false, bool OUT_R = X_MN>
__global__ void __launch_bounds__(THREADS, 1)
    gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmF,
                       const __grid_constant__ CUtensorMap tmF2, const __grid_constant__ CUtensorMap tmD, const Params p) {
  static_assert(!STREAM_F || RSTAGES == ASTAGES, "streamed-F mode indexes both rings with one counter");
  // OUT_R: the output is contiguous along r (the TMEM lanes); else along c.  Resident-factor mode: OUT_R == X_MN.
  constexpr uint32_t XSTRIDE = STREAM_F ? 3 * XT_BYTES : XT_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem - smem_u32(smem_raw));  // generic pointer to the aligned base
  const uint32_t sFhi = smem, sFlo = smem + F_BYTES, sX = STREAM_F ? smem : smem + 2 * F_BYTES;
  const uint32_t sEpi = sX + RSTAGES * XSTRIDE;           // 4 warps x 2 x 4 KB staging blocks (1024-byte aligned)
  const uint32_t bars = sEpi + 8 * EPI_STAGE_BYTES;
  const uint32_t raw_full = bars, raw_free = raw_full + 8 * RSTAGES;
  const uint32_t a_full = raw_free + 8 * RSTAGES, a_free = a_full + 8 * ASTAGES;
  const uint32_t tmem_full = a_free + 8 * ASTAGES, tmem_empty = tmem_full + 16, f_full = tmem_empty + 16;
  const uint32_t epi_bar = f_full + 8;                     // [4 warps][2 buffers]: old-D block has landed
  const uint32_t tmem_slot = epi_bar + 64;
  const uint32_t scratch = tmem_slot + 16;                 // one word per warp (mbar_arrive_after_loads)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < RSTAGES; s++) {
      mbar_init(raw_full + 8 * s, 1);   // TMA transaction bytes
      mbar_init(raw_free + 8 * s, STREAM_F ? 5 : 4);   // one arrival per splitter warp (+ the MMAs' commit)
    }
    for (int s = 0; s < ASTAGES; s++) {
      mbar_init(a_full + 8 * s, 4);     // one arrival per splitter warp
      mbar_init(a_free + 8 * s, 1);     // tcgen05.commit
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tmem_full + 8 * a, 1);    // tcgen05.commit
      mbar_init(tmem_empty + 8 * a, 4);   // one arrival per epilogue warp
    }
    mbar_init(f_full, 1);
    for (int q = 0; q < 8; q++) mbar_init(epi_bar + 8 * q, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // ---- resident factor: TMA the raw FP32 factor into the F_hi region, then split it in place ----
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmF);
    if (STREAM_F) tma_prefetch_desc(&tmF2);
  }
  if (!STREAM_F) {
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(f_full, F_BYTES);
    // four 16 KB boxes: F_K: {32 k, 128 c} per k-block; else {32 c, 128 k} per c-group
    for (int q = 0; q < 4; q++) tma_load_3d(sFhi + q * 16384, &tmF, f_full, 32 * q, 0, 0);
  }
  mbar_wait(f_full, 0);
  {
    float4* hi = reinterpret_cast<float4*>(smem_gen);
    float4* lo = reinterpret_cast<float4*>(smem_gen + F_BYTES);
    for (int i = threadIdx.x; i < F_BYTES / 16; i += THREADS) {
      float4 h, l;
      split4(hi[i], h, l);
      hi[i] = h;
      lo[i] = l;
    }
  }
  fence_async_smem();
  }
  __syncthreads();

  const int nkb = (p.n + BK - 1) / BK;                 // k-blocks of 32
  const int n_mma = STREAM_F ? 128 : ((p.m + 15) / 16) * 16;   // UMMA N (streamed tiles are zero-filled past the edge)
  const long long tiles_rc = (long long)p.tiles_r * p.tiles_c;
  const uint32_t idesc = make_idesc(false, !F_K, n_mma);   // A comes from TMEM (row = lane, k = column)

  // register re-balancing (whole warpgroups): the control warpgroup gives, the epilogue warpgroup takes
  if (warp < 4) {
  if (STREAM_F) setmaxnreg_dec<CTRL_REGS>();
  if (warp == 0) {
    // ===================== TMA producer: raw X tiles (warp-uniform loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = (int)(tile / tiles_rc);
      const int rem = (int)(tile - (long long)b * tiles_rc);
      const int ct = p.c_fast ? rem % p.tiles_c : rem / p.tiles_r;
      const int r0 = (p.c_fast ? rem / p.tiles_c : rem - ct * p.tiles_r) * TM;
      const int bx = p.x_batched ? b : 0;
      for (int kb = 0; kb < nkb; kb++) {
        mbar_wait_relaxed(raw_free + 8 * stage, phase ^ 1);
        if (elect_one()) {
          const uint32_t dst = sX + stage * XSTRIDE;
          const uint32_t bar = raw_full + 8 * stage;
          mbar_arrive_expect_tx(bar, XSTRIDE);
          if (STREAM_F) {
            // the pre-split operand tiles of this k-block: F_K {32 k, 128 c}; else four boxes {32 c, 32 k}
            if (F_K) {
              tma_load_3d(dst + XT_BYTES, &tmF, bar, kb * BK, ct * 128, 0);
              tma_load_3d(dst + 2 * XT_BYTES, &tmF2, bar, kb * BK, ct * 128, 0);
            } else {
#pragma unroll
              for (int q = 0; q < 4; q++) {
                tma_load_3d(dst + XT_BYTES + q * 4096, &tmF, bar, ct * 128 + 32 * q, kb * BK, 0);
                tma_load_3d(dst + 2 * XT_BYTES + q * 4096, &tmF2, bar, ct * 128 + 32 * q, kb * BK, 0);
              }
            }
          }
          if (X_MN) {
#pragma unroll
            for (int q = 0; q < 4; q++) tma_load_3d(dst + q * 4096, &tmX, bar, r0 + 32 * q, kb * BK, bx);
          } else {
            tma_load_3d(dst, &tmX, bar, kb * BK, r0, bx);
          }
        }
        __syncwarp();
        if (++stage == RSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    // factor descriptors: only the 14-bit start-address field (units of 16 B) changes per k-step
    const uint64_t bhi0 = F_K ? make_desc_k(sFhi) : make_desc_mn(sFhi, 16384);
    const uint64_t blo0 = F_K ? make_desc_k(sFlo) : make_desc_mn(sFlo, 16384);
    constexpr uint32_t KB_INC = F_K ? (16384 >> 4) : (4096 >> 4);   // next k-block of 32
    constexpr uint32_t KS_INC = F_K ? (32 >> 4) : (1024 >> 4);      // next k-step of 8
    int at = 0;
    uint32_t aph = 0;
    int it = 0;      // chunk counter of this CTA: accumulator it & 1
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      uint32_t d_tmem = 0;
      for (int kb = 0; kb < nkb; kb++) {
        const bool chunk_first = STREAM_F ? (kb % CHUNK_KB) == 0 : kb == 0;
        const bool chunk_last = STREAM_F ? ((kb % CHUNK_KB) == CHUNK_KB - 1 || kb == nkb - 1) : kb == nkb - 1;
        const int a = it & 1;
        if (chunk_first) {
          mbar_wait(tmem_empty + 8 * a, (uint32_t)((it >> 1) & 1) ^ 1);   // epilogue drained this accumulator
          tc_fence_after();
          d_tmem = tmem_base + a * 128;
        }
        if (STREAM_F) mbar_wait(raw_full + 8 * at, aph);   // the F_hi / F_lo tiles of this stage have landed
        mbar_wait(a_full + 8 * at, aph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t ahi = tmem_base + TMEM_A0 + at * 64, alo = ahi + 32;
          const int ksteps = min(4, (p.n - kb * BK + 7) / 8);
          const uint32_t sF = sX + at * XSTRIDE + XT_BYTES;   // STREAM_F: this stage's F_hi tile, F_lo follows
          const uint64_t shi = F_K ? make_desc_k(sF) : make_desc_mn(sF, 4096);
          const uint64_t slo = F_K ? make_desc_k(sF + XT_BYTES) : make_desc_mn(sF + XT_BYTES, 4096);
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            if (ks < ksteps) {
              const uint64_t bhi = STREAM_F ? shi + (uint64_t)(ks * KS_INC) : bhi0 + (uint64_t)(kb * KB_INC + ks * KS_INC);
              const uint64_t blo = STREAM_F ? slo + (uint64_t)(ks * KS_INC) : blo0 + (uint64_t)(kb * KB_INC + ks * KS_INC);
              const uint32_t first = (chunk_first && ks == 0) ? 0u : 1u;
              if (p.passes == 3) {
                tc_mma_tf32_ts(d_tmem, alo + ks * 8, bhi, idesc, first);
                tc_mma_tf32_ts(d_tmem, ahi + ks * 8, blo, idesc, 1u);
                tc_mma_tf32_ts(d_tmem, ahi + ks * 8, bhi, idesc, 1u);
              } else {
                tc_mma_tf32_ts(d_tmem, ahi + ks * 8, bhi, idesc, first);
              }
            }
          }
          tc_commit(a_free + 8 * at);                 // TMEM A stage reusable once these MMAs retire
          if (STREAM_F) tc_commit(raw_free + 8 * at);   // ... and so are the F tiles of the shared-memory stage
          if (chunk_last) tc_commit(tmem_full + 8 * a);
        }
        __syncwarp();
        if (chunk_last) it++;
        if (++at == ASTAGES) { at = 0; aph ^= 1; }
      }
    }
  }
  } else if (warp >= 4 && warp < 4 + 4 * NSG) {
    // ===================== splitters: raw X tile (smem) -> hi, lo (TMEM A operand) =====================
    const int grp = (warp - 4) >> 2;         // this group handles k-blocks with (global index % NSG) == grp
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;     // tile row handled by this thread
    long long kbg = 0;                       // global k-block counter of this CTA
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < nkb; kb++, kbg++) {
        if ((int)(kbg % NSG) != grp) continue;
        const int rs = (int)(kbg % RSTAGES), at = (int)(kbg % ASTAGES);
        const uint32_t rph = (uint32_t)((kbg / RSTAGES) & 1), aph = (uint32_t)((kbg / ASTAGES) & 1);
        mbar_wait(raw_full + 8 * rs, rph);
        const uint8_t* src = smem_gen + (sX - smem) + rs * XSTRIDE;
        uint32_t hi[32], lo[32];
        if (X_MN) {
          // [4 r-groups][32 k rows][128 B of 32 r], 32-byte chunks XOR (k & 3): lane reads its r for every k
          const uint8_t* g = src + quarter * 4096 + ((lane & 7) << 2);
#pragma unroll
          for (int k = 0; k < 32; k++) {
            const float x = *reinterpret_cast<const float*>(g + k * 128 + ((((lane >> 3) ^ (k & 3))) << 5));
            hi[k] = __float_as_uint(x) & 0xFFFFE000u;
            lo[k] = __float_as_uint(x - __uint_as_float(hi[k])) & 0xFFFFE000u;
          }
        } else {
          // [128 rows][128 B of 32 k], 16-byte chunks XOR (row & 7): lane reads its whole row
          const uint8_t* g = src + row * 128;
#pragma unroll
          for (int cidx = 0; cidx < 8; cidx++) {
            const float4 x = *reinterpret_cast<const float4*>(g + ((cidx ^ (row & 7)) << 4));
            float4 h, l;
            split4(x, h, l);
            hi[4 * cidx + 0] = __float_as_uint(h.x); hi[4 * cidx + 1] = __float_as_uint(h.y);
            hi[4 * cidx + 2] = __float_as_uint(h.z); hi[4 * cidx + 3] = __float_as_uint(h.w);
            lo[4 * cidx + 0] = __float_as_uint(l.x); lo[4 * cidx + 1] = __float_as_uint(l.y);
            lo[4 * cidx + 2] = __float_as_uint(l.z); lo[4 * cidx + 3] = __float_as_uint(l.w);
          }
        }
        // the raw slot is free as soon as every lane holds its values in registers (lo depends on every load)
        uint32_t digest = 0;
#pragma unroll
        for (int k = 0; k < 32; k++) digest ^= lo[k];
        __syncwarp();
        if (lane == 0) mbar_arrive_after_loads(raw_free + 8 * rs, scratch + 4 * warp, digest);
        // hand hi / lo to the tensor core through TMEM
        mbar_wait(a_free + 8 * at, aph ^ 1);
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + TMEM_A0 + at * 64;
        tc_st32(ta, hi);
        tc_st32(ta + 32, lo);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + 8 * at);
      }
    }
  } else if (warp >= 4 + 4 * NSG) {
    if (STREAM_F) setmaxnreg_inc<EPI_REGS>();
    // ===================== epilogue: TMEM -> registers -> (smem -> TMA store | global), alpha/beta fused =====
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const uint32_t stage_s = sEpi + quarter * 2 * EPI_STAGE_BYTES;     // two buffers of this warp
    uint8_t* stage_g = smem_gen + (stage_s - smem);
    const uint32_t ebar = epi_bar + 16 * quarter;
    uint32_t eph0 = 0, eph1 = 0;                    // phases of the two "old block landed" barriers
    const bool readw = (p.beta != 0.f);
    int it = 0;      // chunk counter, in step with the MMA warp
    const int nchunks = STREAM_F ? (nkb + CHUNK_KB - 1) / CHUNK_KB : 1;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = (int)(tile / tiles_rc);
      const int rem = (int)(tile - (long long)b * tiles_rc);
      const int ct = p.c_fast ? rem % p.tiles_c : rem / p.tiles_r;
      const int cbase = ct * 128;                                   // first output column of this tile
      const int nblk = (min(128, p.m - cbase) + 31) / 32;
      const int rw = (p.c_fast ? rem / p.tiles_c : rem - ct * p.tiles_r) * TM + quarter * 32;   // first row of this warp
      const long long r = rw + lane;
      // beta != 0: the old 32 x 32 block of D is TMA-loaded into the staging buffer it will be stored from
      auto load_old = [&](int blk) {
        if (lane == 0) {
          const int buf = blk & 1;
          tma_store_wait_read();      // the store that last used this buffer has been read out
          mbar_arrive_expect_tx(ebar + 8 * buf, EPI_STAGE_BYTES);
          if (OUT_R) tma_load_3d(stage_s + buf * EPI_STAGE_BYTES, &tmD, ebar + 8 * buf, rw, cbase + blk * 32, b);
          else tma_load_3d(stage_s + buf * EPI_STAGE_BYTES, &tmD, ebar + 8 * buf, cbase + blk * 32, rw, 0);
        }
      };
      if (p.tma_store && readw) load_old(0);        // overlaps the wait for the accumulator
      // streamed mode: drain the chunk accumulators into registers (round-to-nearest adds outside the tensor core);
      // resident mode: one chunk, read block by block straight from TMEM below
      float sum[STREAM_F ? 128 : 1];
      int a = it & 1;
      if (STREAM_F) {
        for (int ch = 0; ch < nchunks; ch++, it++) {
          a = it & 1;
          mbar_wait_relaxed(tmem_full + 8 * a, (uint32_t)((it >> 1) & 1));
          tc_fence_after();
#pragma unroll
          for (int blk = 0; blk < 4; blk++) {
            if (blk < nblk) {
              uint32_t v[32];
              tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + a * 128 + blk * 32, v);
              tc_wait_ld();
#pragma unroll
              for (int c = 0; c < 32; c++)
                sum[(STREAM_F ? blk * 32 + c : 0)] = (ch == 0) ? __uint_as_float(v[c]) : sum[(STREAM_F ? blk * 32 + c : 0)] + __uint_as_float(v[c]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty + 8 * a);
        }
      } else {
        mbar_wait_relaxed(tmem_full + 8 * a, (uint32_t)((it >> 1) & 1));
        tc_fence_after();
      }
      float* drow = p.D + (long long)b * p.d_bs + r * p.d_rs;
      const bool row_ok = r < p.rows;
#pragma unroll
      for (int blk = 0; blk < 4; blk++) {
        if (blk >= nblk) break;
        const int c0 = cbase + blk * 32;
        uint32_t v[32];
        if (STREAM_F) {
#pragma unroll
          for (int c = 0; c < 32; c++) v[c] = __float_as_uint(sum[(STREAM_F ? blk * 32 + c : 0)]);
          if (p.nscale > 0 && row_ok) {
#pragma unroll 4
            for (int c = 0; c < 32; c++) {
              if (c0 + c < p.m) {
                const long long mi = p.rows_are_n ? (long long)(c0 + c) : r, ni = p.rows_are_n ? r : (long long)(c0 + c);
                float x = __uint_as_float(v[c]);
                for (int q = 0; q < p.nscale; q++) x *= __ldg(p.scale_ptr[q] + mi + (p.scale_ld[q] ? ni * p.scale_ld[q] : 0));
                v[c] = __float_as_uint(x);
              }
            }
          }
        } else {
          tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + a * 128 + blk * 32, v);
          tc_wait_ld();
        }
        if (p.tma_store) {
          // stage the block in shared memory in the order the output wants, then ONE bulk tensor store: full
          // 128-byte lines, clipped at the matrix edge, no LSU pressure on this warp
          const int buf = blk & 1;
          uint8_t* sg = stage_g + buf * EPI_STAGE_BYTES;
          if (readw) {
            if (blk + 1 < nblk) load_old(blk + 1);  // prefetch the next old block into the other buffer
            if (buf == 0) { mbar_wait(ebar, eph0); eph0 ^= 1; } else { mbar_wait(ebar + 8, eph1); eph1 ^= 1; }
          } else {
            if (lane == 0) tma_store_wait_read1();  // the store issued two blocks ago has been read out
            __syncwarp();
          }
          if (OUT_R) {
            // output is contiguous along r (the lanes): smem [c][32 r]
#pragma unroll
            for (int c = 0; c < 32; c++) {
              float* q = reinterpret_cast<float*>(sg + c * 128 + lane * 4);
              float o = p.alpha * __uint_as_float(v[c]);
              if (readw) o += p.beta * *q;
              *q = o;
            }
          } else {
            // output is contiguous along c: smem [r][32 c], 16-byte chunks XOR (r & 7) (SWIZZLE_128B)
#pragma unroll
            for (int q = 0; q < 8; q++) {
              float4* dst = reinterpret_cast<float4*>(sg + lane * 128 + ((q ^ (lane & 7)) << 4));
              float4 o = make_float4(p.alpha * __uint_as_float(v[4 * q]), p.alpha * __uint_as_float(v[4 * q + 1]),
                                     p.alpha * __uint_as_float(v[4 * q + 2]), p.alpha * __uint_as_float(v[4 * q + 3]));
              if (readw) {
                const float4 old = *dst;
                o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
              }
              *dst = o;
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {   // bulk groups are per thread: the same lane issues, commits and later waits
            if (OUT_R) tma_store_3d(&tmD, stage_s + buf * EPI_STAGE_BYTES, rw, c0, b);
            else tma_store_3d(&tmD, stage_s + buf * EPI_STAGE_BYTES, c0, rw, 0);
            tma_store_commit();
          }
          __syncwarp();
        } else if (row_ok) {
          if (p.vec_ok && c0 + 32 <= p.m) {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              float4 o = make_float4(p.alpha * __uint_as_float(v[c]), p.alpha * __uint_as_float(v[c + 1]),
                                     p.alpha * __uint_as_float(v[c + 2]), p.alpha * __uint_as_float(v[c + 3]));
              float4* dst = reinterpret_cast<float4*>(drow + c0 + c);
              if (p.beta != 0.f) {
                const float4 old = *dst;
                o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
              }
              *dst = o;
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; c++) {
              if (c0 + c < p.m) {
                float* dst = drow + (long long)(c0 + c) * p.d_cs;
                float o = p.alpha * __uint_as_float(v[c]);
                if (p.beta != 0.f) o += p.beta * *dst;
                *dst = o;
              }
            }
          }
        }
      }
      if (!STREAM_F) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty + 8 * a);
        it++;
      }
    }
    if (lane == 0) tma_store_wait_all();   // this lane's bulk stores are complete before the CTA exits
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// hi / lo halves of an FP32 matrix, column by column: only the `rows` stored entries of each column are read (a view
// with ld > rows may end at its allocation boundary), written with leading dimension ld_out (a multiple of 4).
// VEC: src 16-byte aligned and ld % 4 == 0 -> float4 for the full groups of 4 rows, scalars for the tail.
template <bool VEC>
__global__ void split_hilo_kernel(const float* __restrict__ src, long long ld, long long rows, long long cols,
                                  float* __restrict__ hi, float* __restrict__ lo, long long ld_out) {
  const long long r4 = VEC ? rows / 4 : 0;                  // float4 groups per column
  const long long per_col = r4 + (rows - 4 * r4);           // work items per column: groups, then tail scalars
  const long long total = per_col * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / per_col, q = i - c * per_col;
    if (q < r4) {
      float4 h, l;
      split4(__ldcs(reinterpret_cast<const float4*>(src + c * ld) + q), h, l);
      reinterpret_cast<float4*>(hi + c * ld_out)[q] = h;
      reinterpret_cast<float4*>(lo + c * ld_out)[q] = l;
    } else {
      const long long r = 4 * r4 + (q - r4);
      const float x = src[c * ld + r];
      const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
      hi[c * ld_out + r] = h;
      lo[c * ld_out + r] = __uint_as_float(__float_as_uint(x - h) & 0xFFFFE000u);
    }
  }
}

}  // namespace tf32
}  // namespace scimlop
