// gemm_dmma.cuh -- FP64 strided-batched GEMM for sm_100a: TMA-staged operand tiles (128B-swizzled),
// mbarrier producer/consumer ring, DMMA (mma.sync.m8n8k4.f64) math, fused epilogue.
//
//   C_b(M x N) = alpha * rowscale .* (A_b(M x K) * B_b(K x N)) + beta * C_b          (column-major C)
//
// This is the engine behind every contraction of the hot path: the reshape /
// permutedims! of src/tensor.jl:558,780-787 become pure tensor-map geometry (an operand is either
// "outer-contiguous" or "k-contiguous" in global memory and TMA lays it out in shared memory for us),
// and copy!/axpby! (src/tensor.jl:829-831), the ScaledOperator λ (src/basic.jl:518-536) and a left
// DiagonalOperator / BatchedDiagonalOperator (src/basic.jl:1160-1189) live in the epilogue.
//
// tcgen05 has no FP64 kind, so FP64 tensor math on Blackwell is the warp-level DMMA.8x8x4 (all larger
// mma.sync f64 shapes lower to it; measured peak 37.1 TFLOP/s, profiles/r01_peaks_microbench.json).
#pragma once
#include "ptx.cuh"

namespace scimlop {
namespace dmma {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 5;
constexpr int TILE_BYTES = BM * BK * 8;      // 16 KB per operand tile
constexpr int STAGE_BYTES = 2 * TILE_BYTES;  // A + B
constexpr int CONSUMER_WARPS = 8;            // 2 (m) x 4 (n), warp tile 64 x 32
constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;    // + one producer warpgroup
constexpr int CONSUMER_REGS = 232, PRODUCER_REGS = 40;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int MAX_SCALES = 3;

struct Scale {
  const double* ptr;
  long long ld;  // 0: vector indexed by row m; else matrix indexed m + n*ld
};

struct Params {
  int M, N, K, batch;
  int tiles_m, tiles_n;
  long long total_tiles;   // tiles THIS launch covers: global tile ids tile_base .. tile_base + total_tiles - 1
  int tile_base;           // != 0 for the split-K launch that takes the last partial wave of a many-tile problem
  double* C;
  long long ldc, strideC;
  double alpha, beta;  // beta == 0: C is never read
  int nscale;
  Scale scale[MAX_SCALES];
  int vec_ok;  // C base, ldc, strideC allow 16-byte stores
  int a_batched, b_batched;
  // Fused "compute + all-gather": every output tile is ALSO stored into the same place of up to 7 peer GPUs'
  // copies of C (peer-mapped pointers, NVLink stores from the epilogue), so a replicated result needs no
  // separate collective.  Only with beta == 0.
  int npeer;
  double* peerC[7];
  // NVSwitch multicast alias of C (same geometry): when set, every output element is stored ONCE with multimem.st
  // and the switch delivers it to every rank's copy, this one included.
  double* mcC;
};

__device__ __forceinline__ void multimem_st16(double* addr, double a, double b) {
  const long long x = __double_as_longlong(a), y = __double_as_longlong(b);
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(__int_as_float((int)x)),
               "f"(__int_as_float((int)(x >> 32))), "f"(__int_as_float((int)y)), "f"(__int_as_float((int)(y >> 32)))
               : "memory");
}
__device__ __forceinline__ void multimem_st8(double* addr, double a) {
  asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(addr), "d"(a) : "memory");
}

// k index used by lane t in MMA step s of a BK=16 block.  The permutation makes the 64-bit fragment
// loads bank-conflict-free for BOTH swizzled tile layouts (see tile_off): the four k's of a step hit
// four different {2c,2c+1} pairs mod 8 (outer-contiguous tiles) and come as two even + two odd values
// whose halves differ in bit 2 (k-contiguous tiles).  Any permutation of k is legal for a dot product.
__device__ __forceinline__ int kperm(int s, int t) {
  // s0 {0,10,5,15}  s1 {1,11,4,14}  s2 {2,8,7,13}  s3 {3,9,6,12}   (one nibble per t)
  const uint32_t tab = (s == 0) ? 0xF5A0u : (s == 1) ? 0xE4B1u : (s == 2) ? 0xD782u : 0xC693u;
  return (int)((tab >> (4 * t)) & 15u);
}

// Byte offset of element (o, k) inside one 16 KB operand tile.
//  KC (k-contiguous source):   one TMA box {16 k, 128 outer}: row o = 128 B, 16-byte chunks XOR (o & 7).
//  OC (outer-contiguous source): eight TMA boxes {16 outer, 16 k} of 2 KB: row k = 128 B, chunks XOR (k & 7).
template <bool KC>
__device__ __forceinline__ uint32_t tile_off(int o, int k) {
  if (KC) return o * 128 + ((((k >> 1) ^ (o & 7))) << 4) + ((k & 1) << 3);
  return (o >> 4) * 2048 + k * 128 + (((((o & 15) >> 1) ^ (k & 7))) << 4) + ((o & 1) << 3);
}

// Warp-specialised: warps 0-7 (two warpgroups) are DMMA consumers, warpgroup 2 (warps 8-11) is the TMA
// producer (one elected lane).  An SM sub-partition owns 16K registers, so three resident warps per
// sub-partition launch at <= 168 registers each; the producer group then drops to PRODUCER_REGS with
// setmaxnreg.dec and the consumers grow to CONSUMER_REGS with setmaxnreg.inc (2*232 + 40 <= 512), which is
// what lets a 64x32 FP64 warp tile (128 accumulator registers) plus double-buffered fragments live in
// registers without spilling.
//
// SPLIT = true is the few-tiles form (a vector-input Kronecker product, a small MatrixOperator): one tile per
// thread-block CLUSTER, the K range divided over the cluster's CTAs, partial tiles summed through distributed
// shared memory in rank order (each rank reduces and stores 1/size of the tile).  Launched with
// grid = cluster size x tiles.
// PEER = true is the fused compute + all-gather form with BULK stores: the ring shrinks to 3 stages and the freed
// shared memory holds one finished 128 x 128 tile, which ONE thread then sends to the local C and to every peer's
// C with a bulk tensor store each (cp.async.bulk.tensor shared -> global over NVLink: full-size packets, no LSU
// traffic, asynchronous to the next tile's DMMAs).  Needs beta == 0, no row scalings, TMA-addressable C.
constexpr int PEER_STAGES = 3;
constexpr int PEER_SMEM_BYTES = PEER_STAGES * STAGE_BYTES + BM * BN * 8 + 1024 + 256;
struct PeerMaps {
  CUtensorMap m[8];   // [0] = local C, [1..npeer] = the peers' C (same geometry)
};
constexpr int PART_PITCH = BM + 8;   // doubles per column of the partial tile in shared memory (conflict-free stores)
static_assert(PART_PITCH * BN * 8 <= STAGES * STAGE_BYTES, "partial tile must fit in the (idle) operand ring");

template <bool A_KC, bool B_OC, bool SPLIT = false, bool PEER = false>
__global__ void __launch_bounds__(THREADS, 1)
    gemm_dmma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const Params p, const __grid_constant__ PeerMaps pm) {
  static_assert(!(SPLIT && PEER), "split-K and peer stores are separate launch modes");
  constexpr int NSTAGE = PEER ? PEER_STAGES : STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 128B swizzle patterns repeat every 1024 B: align the tile ring (in the 32-bit shared window).
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t out_tile = smem + NSTAGE * STAGE_BYTES;   // PEER: the finished tile, dense [n][128 m]
  const uint32_t full_bar = out_tile + (PEER ? BM * BN * 8 : 0);  // NSTAGE x 8 B
  const uint32_t empty_bar = full_bar + NSTAGE * 8;
  // PEER + multicast: out_tile hand-over between the consumers and the three copy warps of the producer warpgroup
  const uint32_t tile_full = full_bar + 64, tile_empty = full_bar + 72;
  const uint32_t scratch = full_bar + 128;  // one word per consumer warp (mbar_arrive_after_loads)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (PEER) {
      mbar_init(tile_full, 1);
      mbar_init(tile_empty, 3);
    }
    for (int s = 0; s < NSTAGE; s++) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, CONSUMER_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const int nk_all = (p.K + BK - 1) / BK;
  const int ntiles = (int)p.total_tiles;
  // SPLIT: this CTA's share of the k-blocks of its cluster's single tile
  const int crank = SPLIT ? (int)cluster_rank() : 0, csize = SPLIT ? (int)cluster_size() : 1;
  const int kb_first = SPLIT ? (int)((long long)nk_all * crank / csize) : 0;
  const int nk = SPLIT ? (int)((long long)nk_all * (crank + 1) / csize) - kb_first : nk_all;
  const int tile_first = SPLIT ? (int)(blockIdx.x / csize) : (int)blockIdx.x;
  const int tile_step = SPLIT ? ntiles : (int)gridDim.x;   // SPLIT: exactly one tile per cluster

  if (warp >= CONSUMER_WARPS) {
    // ================= TMA producer warpgroup =================
    setmaxnreg_dec<PRODUCER_REGS>();
    if (warp == CONSUMER_WARPS && lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      const int tiles_mn = p.tiles_m * p.tiles_n;
      for (int tile = tile_first; tile < ntiles; tile += tile_step) {
        const int b = (tile + p.tile_base) / tiles_mn;
        const int rem = (tile + p.tile_base) - b * tiles_mn;
        const int nt = rem / p.tiles_m;
        const int mt = rem - nt * p.tiles_m;
        const int m0 = mt * BM, n0 = nt * BN;
        const int ba = p.a_batched ? b : 0, bb = p.b_batched ? b : 0;
        for (int kb = kb_first; kb < kb_first + nk; kb++) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t sA = smem + stage * STAGE_BYTES;
          const uint32_t sB = sA + TILE_BYTES;
          const uint32_t bar = full_bar + 8 * stage;
          mbar_arrive_expect_tx(bar, STAGE_BYTES);
          const int k0 = kb * BK;
          if (A_KC) {
            tma_load_3d(sA, &tmA, bar, k0, m0, ba);
          } else {
#pragma unroll
            for (int q = 0; q < 8; q++) tma_load_3d(sA + q * 2048, &tmA, bar, m0 + 16 * q, k0, ba);
          }
          if (!B_OC) {
            tma_load_3d(sB, &tmB, bar, k0, n0, bb);
          } else {
#pragma unroll
            for (int q = 0; q < 8; q++) tma_load_3d(sB + q * 2048, &tmB, bar, n0 + 16 * q, k0, bb);
          }
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (PEER && p.mcC != nullptr && warp > CONSUMER_WARPS) {
      // ---- copy warps (multicast form): finished tile, shared memory -> multimem.st, off the consumers' path ----
      const int ct = (warp - CONSUMER_WARPS - 1) * 32 + lane;      // 0..95
      const int tiles_mn = p.tiles_m * p.tiles_n;
      uint32_t tph = 0;
      for (int tile = tile_first; tile < ntiles; tile += tile_step) {
        const int b = (tile + p.tile_base) / tiles_mn;
        const int rem = (tile + p.tile_base) - b * tiles_mn;
        const int nt = rem / p.tiles_m;
        const int mt = rem - nt * p.tiles_m;
        double* base = p.mcC + (long long)b * p.strideC + (long long)(nt * BN) * p.ldc + mt * BM;
        const int mlim = p.M - mt * BM, nlim = p.N - nt * BN;     // valid extent of this tile
        mbar_wait(tile_full, tph);
        for (int c = ct; c < BM * BN / 2; c += 96) {               // 16-byte pieces: two consecutive rows of a column
          const int n = c / (BM / 2), m = 2 * (c % (BM / 2));
          if (n < nlim && m < mlim) {
            double v0, v1;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(out_tile + (uint32_t)c * 16u));
            if (m + 1 < mlim) multimem_st16(base + (long long)n * p.ldc + m, v0, v1);
            else multimem_st8(base + (long long)n * p.ldc + m, v0);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(tile_empty);
        tph ^= 1;
      }
      __threadfence_system();
    }
    if (SPLIT) {   // the whole CTA takes part in the cluster barriers of the reduction
      __syncwarp();
      cluster_sync_all();
      cluster_sync_all();
    }
    return;
  }

  // ================= DMMA consumers =================
  setmaxnreg_inc<CONSUMER_REGS>();
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1;   // 64-row half of the tile
  const int wn = warp >> 1;  // 32-column quarter of the tile

  // per-lane fragment offsets for the four MMA steps of a stage (sub-tile 0 and 1; the rest are immediates)
  uint32_t offA[4][2], offB[4][2];
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int k = kperm(s, t);
    offA[s][0] = smem + tile_off<A_KC>(wm * 64 + g, k);
    offA[s][1] = smem + tile_off<A_KC>(wm * 64 + 8 + g, k);
    offB[s][0] = smem + TILE_BYTES + tile_off<!B_OC>(wn * 32 + g, k);
    offB[s][1] = smem + TILE_BYTES + tile_off<!B_OC>(wn * 32 + 8 + g, k);
  }
  constexpr uint32_t PAIR = 2048;  // +16 outer rows in either layout

  // fragments of MMA step s of the stage at byte offset `so`
#define SCIMLOP_LOAD_FRAGS(A_, B_, so, s)                                               \
  {                                                                                     \
    _Pragma("unroll") for (int i = 0; i < 8; i++)                                       \
        A_[i] = lds_f64(offA[s][i & 1] + (so) + (i >> 1) * PAIR);                       \
    _Pragma("unroll") for (int j = 0; j < 4; j++)                                       \
        B_[j] = lds_f64(offB[s][j & 1] + (so) + (j >> 1) * PAIR);                       \
  }
  // MMA rows <- n (B operand), MMA cols <- m (A operand): lane owns C[m = 2t, 2t+1][n = g]
#define SCIMLOP_MMA_STEP(A_, B_)                                                        \
  {                                                                                     \
    _Pragma("unroll") for (int i = 0; i < 8; i++)                                       \
        _Pragma("unroll") for (int j = 0; j < 4; j++)                                   \
            dmma884(acc[i][j][0], acc[i][j][1], B_[j], A_[i]);                          \
  }

  int stage = 0;
  uint32_t phase = 0, out_phase = 0;
  double a0[8], b0[4], a1[8], b1[4];  // double-buffered fragments: step s computes while step s+1 loads

  // prime the software pipeline with the first fragments of this CTA's first stage
  if (tile_first < ntiles && nk > 0) {
    mbar_wait(full_bar, 0);
    SCIMLOP_LOAD_FRAGS(a0, b0, 0u, 0);
  }

  for (int tile = tile_first; tile < ntiles; tile += tile_step) {
    const int tiles_mn = p.tiles_m * p.tiles_n;
    const int b = (tile + p.tile_base) / tiles_mn;
    const int rem = (tile + p.tile_base) - b * tiles_mn;
    const int nt = rem / p.tiles_m;
    const int mt = rem - nt * p.tiles_m;
    const bool last_tile = tile + tile_step >= ntiles;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kb = 0; kb < nk; kb++) {
      const uint32_t so = stage * STAGE_BYTES;
      SCIMLOP_LOAD_FRAGS(a1, b1, so, 1);
      SCIMLOP_MMA_STEP(a0, b0);
      SCIMLOP_LOAD_FRAGS(a0, b0, so, 2);
      SCIMLOP_MMA_STEP(a1, b1);
      SCIMLOP_LOAD_FRAGS(a1, b1, so, 3);
      SCIMLOP_MMA_STEP(a0, b0);
      // prefetch step 0 of the next stage (next k-block, or the first k-block of this CTA's next tile)
      const uint32_t done_bar = empty_bar + 8 * stage;
      if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      if (!(last_tile && kb == nk - 1)) {
        mbar_wait(full_bar + 8 * stage, phase);
        SCIMLOP_LOAD_FRAGS(a0, b0, stage * STAGE_BYTES, 0);
      }
      SCIMLOP_MMA_STEP(a1, b1);
      // hand the finished slot back to the producer once its fragments have ARRIVED in registers (see
      // mbar_arrive_after_loads).  Steps 1-2 were consumed before a0/b0 and a1/b1 were overwritten above; the
      // digest covers step 3.  Placed after the last MMA step so that waiting for the data costs no issue slots.
      {
        uint32_t digest = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) digest ^= fold_bits(a1[i]);
#pragma unroll
        for (int j = 0; j < 4; j++) digest ^= fold_bits(b1[j]);
        __syncwarp();
        if (lane == 0) mbar_arrive_after_loads(done_bar, scratch + 4 * warp, digest);
      }
    }
#undef SCIMLOP_LOAD_FRAGS
#undef SCIMLOP_MMA_STEP

    if (SPLIT) {
      // ---- split-K epilogue: partial tile -> shared memory (the operand ring is idle now), cluster barrier,
      //      every rank sums its share of the columns over all ranks (rank order) and runs the fused epilogue ----
      asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory");   // all consumer warps are done reading the ring
      {
        const uint32_t pbase = smem + (uint32_t)(((wn * 32 + g) * PART_PITCH + wm * 64 + 2 * t) * 8);
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int i = 0; i < 8; i++)
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(pbase + (uint32_t)((8 * j * PART_PITCH + 8 * i) * 8)),
                         "d"(acc[i][j][0]), "d"(acc[i][j][1])
                         : "memory");
      }
      cluster_sync_all();
      // rank r owns tile columns [r*ncols, ...): two rows (16 bytes) per thread and step, remote loads of all ranks
      // issued before the first use
      const int ncols = (BN + csize - 1) / csize;
      const int n_lo = crank * ncols, n_hi = min(BN, n_lo + ncols);
      for (int e = (int)threadIdx.x; e < (n_hi - n_lo) * (BM / 2); e += CONSUMER_WARPS * 32) {
        const int nl = n_lo + e / (BM / 2), ml = 2 * (e % (BM / 2));
        const int m = mt * BM + ml, n = nt * BN + nl;
        if (m >= p.M || n >= p.N) continue;
        const uint32_t off = smem + (uint32_t)((nl * PART_PITCH + ml) * 8);
        double p0[8], p1[8];
#pragma unroll
        for (int s = 0; s < 8; s++) {
          p0[s] = p1[s] = 0.0;
          if (s < csize)
            asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(p0[s]), "=d"(p1[s]) : "r"(map_to_rank(off, s)));
        }
        double v0 = p0[0], v1 = p1[0];
#pragma unroll
        for (int s = 1; s < 8; s++) { v0 += p0[s]; v1 += p1[s]; }
        const bool pair = m + 1 < p.M;
        for (int q = 0; q < p.nscale; q++) {
          const double* sp = p.scale[q].ptr + (p.scale[q].ld ? (long long)n * p.scale[q].ld : 0) + m;
          v0 *= sp[0];
          if (pair) v1 *= sp[1];
        }
        double* dst = p.C + (long long)b * p.strideC + (long long)n * p.ldc + m;
        if (pair && p.vec_ok) {
          double2 o = make_double2(p.alpha * v0, p.alpha * v1);
          if (p.beta != 0.0) {
            const double2 old = *reinterpret_cast<const double2*>(dst);
            o.x = fma(p.beta, old.x, o.x);
            o.y = fma(p.beta, old.y, o.y);
          }
          *reinterpret_cast<double2*>(dst) = o;
        } else {
          dst[0] = (p.beta == 0.0) ? p.alpha * v0 : fma(p.alpha, v0, p.beta * dst[0]);
          if (pair) dst[1] = (p.beta == 0.0) ? p.alpha * v1 : fma(p.alpha, v1, p.beta * dst[1]);
        }
      }
      cluster_sync_all();   // nobody leaves while a peer may still read its partial tile
      return;
    }
    if (PEER) {
      // ---- bulk-store epilogue: tile -> shared memory -> one bulk tensor store per destination ----
      const bool mc = p.mcC != nullptr;
      if (threadIdx.x == 0) {                           // out_tile is free again:
        if (mc) mbar_wait(tile_empty, out_phase ^ 1);   //   the copy warps have sent the previous tile
        else tma_store_wait_read();                     //   the previous tile's bulk stores have been read out
      }
      asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory");
      {
        const double alpha = p.alpha;
        const uint32_t obase = out_tile + (uint32_t)(((wn * 32 + g) * BM + wm * 64 + 2 * t) * 8);
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int i = 0; i < 8; i++)
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(obase + (uint32_t)((8 * j * BM + 8 * i) * 8)),
                         "d"(alpha * acc[i][j][0]), "d"(alpha * acc[i][j][1])
                         : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory");
      if (threadIdx.x == 0) {
        if (mc) {
          mbar_arrive(tile_full);                          // (release: the tile is complete in shared memory)
        } else {
          for (int q = 0; q <= p.npeer; q++) tma_store_3d(&pm.m[q], out_tile, mt * BM, nt * BN, b);
          tma_store_commit();
          if (last_tile) tma_store_wait_all();              // everything has landed before the CTA retires
        }
      }
      out_phase ^= 1;
      continue;
    }
    // ---- fused epilogue: rowscale, alpha, beta, store (the next tile's first stages are already in flight) ----
    const int mbase = mt * BM + wm * 64 + 2 * t;
    const int nbase = nt * BN + wn * 32 + g;
    double* Cb = p.C + (long long)b * p.strideC + (long long)nbase * p.ldc + mbase;
    const bool interior = p.vec_ok && (mt + 1) * BM <= p.M && (nt + 1) * BN <= p.N;
    if (interior && p.nscale == 0) {
      // fast path (every tile of the headline shapes): 32 x 16-byte stores per lane, no bounds checks
      const double alpha = p.alpha, beta = p.beta;
      if (beta == 0.0 && p.mcC != nullptr) {
        double* Mb = p.mcC + (Cb - p.C);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          double* col = Mb + (long long)(8 * j) * p.ldc;
#pragma unroll
          for (int i = 0; i < 8; i++) multimem_st16(col + 8 * i, alpha * acc[i][j][0], alpha * acc[i][j][1]);
        }
      } else if (beta == 0.0) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          double* col = Cb + (long long)(8 * j) * p.ldc;
#pragma unroll
          for (int i = 0; i < 8; i++)
            *reinterpret_cast<double2*>(col + 8 * i) = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
        }
        if (p.npeer > 0) {
          // the same tile into every peer's C (peer-mapped memory: these stores travel over NVLink)
          const long long off = Cb - p.C;
          for (int q = 0; q < p.npeer; q++) {
            double* Pb = p.peerC[q] + off;
#pragma unroll
            for (int j = 0; j < 4; j++) {
              double* col = Pb + (long long)(8 * j) * p.ldc;
#pragma unroll
              for (int i = 0; i < 8; i++)
                *reinterpret_cast<double2*>(col + 8 * i) = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          double* col = Cb + (long long)(8 * j) * p.ldc;
          double2 old[8];
#pragma unroll
          for (int i = 0; i < 8; i++) old[i] = *reinterpret_cast<const double2*>(col + 8 * i);
#pragma unroll
          for (int i = 0; i < 8; i++)
            *reinterpret_cast<double2*>(col + 8 * i) =
                make_double2(fma(alpha, acc[i][j][0], beta * old[i].x), fma(alpha, acc[i][j][1], beta * old[i].y));
        }
      }
    } else {
      // general path: edge tiles, unaligned C, diagonal row scalings
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int n = nbase + 8 * j;
        if (n >= p.N) continue;
        double* col = Cb + (long long)(8 * j) * p.ldc;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int m = mbase + 8 * i;
          if (m >= p.M) continue;
          double v0 = acc[i][j][0], v1 = acc[i][j][1];
          const bool pair = (m + 1 < p.M);
          for (int q = 0; q < p.nscale; q++) {
            const double* sp = p.scale[q].ptr + (p.scale[q].ld ? (long long)n * p.scale[q].ld : 0) + m;
            v0 *= sp[0];
            if (pair) v1 *= sp[1];
          }
          v0 *= p.alpha;
          v1 *= p.alpha;
          double* dst = col + 8 * i;
          if (p.mcC != nullptr) {      // (beta == 0 and no scalings are enforced on the host)
            double* md = p.mcC + (dst - p.C);
            multimem_st8(md, v0);
            if (pair) multimem_st8(md + 1, v1);
            continue;
          }
          if (pair && p.vec_ok) {
            if (p.beta != 0.0) {
              const double2 old = *reinterpret_cast<const double2*>(dst);
              v0 += p.beta * old.x;
              v1 += p.beta * old.y;
            }
            *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
          } else {
            if (p.beta != 0.0) {
              v0 += p.beta * dst[0];
              if (pair) v1 += p.beta * dst[1];
            }
            dst[0] = v0;
            if (pair) dst[1] = v1;
          }
          for (int q = 0; q < p.npeer; q++) {   // (beta == 0 is enforced on the host when peers are given)
            double* pd = p.peerC[q] + (dst - p.C);
            pd[0] = v0;
            if (pair) pd[1] = v1;
          }
        }
      }
    }
  }
  if (p.mcC != nullptr) __threadfence_system();   // multicast stores are performed before the kernel retires
}

}  // namespace dmma
}  // namespace scimlop
