// scimlop.cu -- C ABI of libscimlop_b200.so: handle, GEMM dispatch, elementwise pass, and the host-side
// planners that turn a TensorProductOperator / operator tree into kernel launches.
//
// Reference behaviour being replaced (SciMLOperators.jl): src/tensor.jl:543-610,750-834 (TPO mul! and
// outer_mul!), src/matrix.jl:337-361, src/batch.jl:151-179, src/basic.jl:74-90,248-264,518-536,834-854,
// 1160-1207, src/block.jl:145-179.  No CPU fallback: every path ends in a CUDA kernel launch or an error status.
//
// GEMM dispatch (gemm_dispatch<T>), in order of preference:
//   Float64: few columns -> gemm_skinny (TMA ring, cluster/DSMEM reduce, HBM-bound); aligned -> gemm_dmma (FP64
//            DMMA; split-K over clusters when tiles < SMs; PEER variant for the fused all-gather); else gemm_simt.
//   Float32: small resident factor -> gemm_tf32x3 (tcgen05, TMEM); few columns -> gemm_skinny; large ->
//            gemm_tf32x3<STREAM_F> (pre-split operand streamed, chunked accumulation); else / FP32_EXACT -> gemm_simt.
#include <algorithm>
#include <type_traits>
#include <new>

#include "context.h"
#include "banded.cuh"
#include "eltwise.cuh"
#include "gemm_dmma.cuh"
#include "gemm_simt.cuh"
#include "gemm_skinny.cuh"
#include "gemm_tf32.cuh"
#include "kron2_f16.cuh"
#include "kron2_f64_small.cuh"

using namespace scimlop;

// ================================================================================================
// small helpers
// ================================================================================================
namespace {

template <typename T>
T scalar_or(const void* p, T dflt) {
  return p ? *static_cast<const T*>(p) : dflt;
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline size_t elem_size(int dtype) { return dtype == SCIMLOP_F64 ? 8 : 4; }

struct ScaleArg {
  const void* ptr;
  long long ld;
};
// peer destinations of a fused compute + all-gather launch: passed explicitly down to the LAST kron stage
struct PeerArg {
  int n = 0;
  void* ptr[7] = {};
  void* mc = nullptr;   // multicast alias of the destination (scimlop_kron_mul_mc): one store reaches every rank
  bool active() const { return n > 0 || mc != nullptr; }
};
const PeerArg NO_PEERS;

// ---- elementwise launcher -------------------------------------------------------------------------
template <typename T>
int launch_eltwise(scimlop_context* c, elt::Params<T>& p, cudaStream_t st) {
  if (p.n <= 0 || p.k <= 0) return SCIMLOP_OK;
  constexpr int VMAX = (sizeof(T) == 8) ? 2 : 4;
  bool vec = (p.n % VMAX == 0) && (p.ldw % VMAX == 0) && aligned16(p.W);
  if (p.nterms > 0) vec = vec && (p.ldv % VMAX == 0) && aligned16(p.V);
  for (int t = 0; t < p.nterms; t++)
    for (int f = p.nvec[t]; f < p.nfac[t]; f++)
      vec = vec && (p.fac[t][f].ld % VMAX == 0) && aligned16(p.fac[t][f].ptr);
  const int V = vec ? VMAX : 1;
  const long long rows = (p.n + V - 1) / V;
  const long long row_blocks = (rows + elt::THREADS - 1) / elt::THREADS;
  // enough blocks to fill the machine several times over, each looping over a multiple of UNROLL columns
  const long long target = (long long)c->sm_count * 16;
  long long col_blocks = std::max<long long>(1, std::min<long long>((target + row_blocks - 1) / row_blocks, 65535));
  long long cpb = (p.k + col_blocks - 1) / col_blocks;
  constexpr long long CU = 8;  // column unroll of the widest kernel
  cpb = std::max<long long>(CU, (cpb + CU - 1) / CU * CU);
  col_blocks = (p.k + cpb - 1) / cpb;
  if (col_blocks > 65535) {
    cpb = ((p.k + 65534) / 65535 + CU - 1) / CU * CU;
    col_blocks = (p.k + cpb - 1) / cpb;
  }
  if (row_blocks > 2147483647LL) return scimlop_fail(c, SCIMLOP_ERR_BAD_SHAPE, "eltwise: too many rows");
  p.cols_per_block = (int)cpb;
  dim3 grid((unsigned)row_blocks, (unsigned)col_blocks);
  bool batched = false;
  for (int t = 0; t < p.nterms; t++) batched = batched || (p.nfac[t] > p.nvec[t]);
  if (!batched) {
    if (vec) elt::eltwise_rowcoef_kernel<T, VMAX><<<grid, elt::THREADS, 0, st>>>(p);
    else elt::eltwise_rowcoef_kernel<T, 1><<<grid, elt::THREADS, 0, st>>>(p);
  } else if (vec)
    elt::eltwise_kernel<T, VMAX><<<grid, elt::THREADS, 0, st>>>(p);
  else
    elt::eltwise_kernel<T, 1><<<grid, elt::THREADS, 0, st>>>(p);
  c->launches++;
  SCIMLOP_CUDA(c, cudaGetLastError());
  return SCIMLOP_OK;
}

// W = alpha * V + beta * W  (V may be NULL with alpha ignored: W = beta * W)
template <typename T>
int eltwise_axpby(scimlop_context* c, long long n, long long k, const T* V, long long ldv, T* W, long long ldw,
                  T alpha, T beta, cudaStream_t st) {
  elt::Params<T> p;
  memset(&p, 0, sizeof(p));
  p.n = n; p.k = k; p.V = V; p.ldv = ldv; p.W = W; p.ldw = ldw; p.alpha = alpha; p.beta = beta;
  if (V) { p.nterms = 1; p.coef[0] = T(1); }
  return launch_eltwise(c, p, st);
}

// ---- DMMA (FP64, TMA) launcher ------------------------------------------------------------------------
int make_map_f64(scimlop_context* c, CUtensorMap* map, const double* base, bool kc, long long outer, long long K,
                 long long ld, long long stride, long long nb) {
  cuuint64_t gdim[3];
  cuuint64_t gstr[2];
  cuuint32_t box[3];
  cuuint32_t estr[3] = {1, 1, 1};
  if (kc) {
    gdim[0] = (cuuint64_t)K; gdim[1] = (cuuint64_t)outer;
    box[0] = dmma::BK; box[1] = dmma::BM;
  } else {
    gdim[0] = (cuuint64_t)outer; gdim[1] = (cuuint64_t)K;
    box[0] = 16; box[1] = dmma::BK;
  }
  gdim[2] = (cuuint64_t)nb;
  box[2] = 1;
  gstr[0] = (cuuint64_t)ld * 8;
  gstr[1] = (nb > 1) ? (cuuint64_t)stride * 8 : (cuuint64_t)ld * 8 * gdim[1];
  CUresult r = c->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), gdim, gstr, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return scimlop_fail(c, SCIMLOP_ERR_CUDA,
                        "cuTensorMapEncodeTiled failed (CUresult %d): dims {%llu,%llu,%llu} strides {%llu,%llu}",
                        (int)r, (unsigned long long)gdim[0], (unsigned long long)gdim[1],
                        (unsigned long long)gdim[2], (unsigned long long)gstr[0], (unsigned long long)gstr[1]);
  return SCIMLOP_OK;
}

bool dmma_eligible(long long M, long long N, long long K, const double* A, long long lda, long long sA,
                   const double* B, long long ldb, long long sB, long long batch) {
  if (M < 16 || N < 16 || K < 4) return false;
  if (!aligned16(A) || !aligned16(B)) return false;
  if ((lda & 1) || (ldb & 1)) return false;
  if (batch > 1 && ((sA & 1) || (sB & 1))) return false;
  const long long lim = 2147483647LL;
  if (M > lim || N > lim || K > lim || batch > lim) return false;
  if (lda > (1LL << 36) || ldb > (1LL << 36) || sA > (1LL << 36) || sB > (1LL << 36)) return false;
  return true;
}

int launch_dmma(scimlop_context* c, bool transA, bool transB, long long M, long long N, long long K, double alpha,
                const double* A, long long lda, long long sA, const double* B, long long ldb, long long sB,
                double beta, double* C, long long ldc, long long sC, long long batch, int nscale,
                const ScaleArg* scales, cudaStream_t st, const PeerArg& g_peers = NO_PEERS) {
  const bool a_batched = batch > 1 && sA != 0, b_batched = batch > 1 && sB != 0;
  CUtensorMap tmA, tmB;
  int rc = make_map_f64(c, &tmA, A, /*kc=*/transA, M, K, lda, sA, a_batched ? batch : 1);
  if (rc) return rc;
  rc = make_map_f64(c, &tmB, B, /*kc=*/!transB, N, K, ldb, sB, b_batched ? batch : 1);
  if (rc) return rc;
  dmma::Params p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.batch = (int)batch;
  p.tiles_m = (int)((M + dmma::BM - 1) / dmma::BM);
  p.tiles_n = (int)((N + dmma::BN - 1) / dmma::BN);
  p.total_tiles = (long long)p.tiles_m * p.tiles_n * batch;
  p.C = C; p.ldc = ldc; p.strideC = sC;
  p.alpha = alpha; p.beta = beta;
  p.nscale = nscale;
  for (int q = 0; q < nscale; q++) { p.scale[q].ptr = static_cast<const double*>(scales[q].ptr); p.scale[q].ld = scales[q].ld; }
  p.vec_ok = aligned16(C) && !(ldc & 1) && (batch == 1 || !(sC & 1));
  p.a_batched = a_batched; p.b_batched = b_batched;
  p.npeer = g_peers.n;
  for (int q = 0; q < g_peers.n; q++) {
    p.peerC[q] = static_cast<double*>(g_peers.ptr[q]);
    p.vec_ok = p.vec_ok && aligned16(g_peers.ptr[q]);
  }
  p.mcC = static_cast<double*>(g_peers.mc);
  if (g_peers.active() && (beta != 0.0 || nscale != 0))
    return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "peer / multicast targets need beta == 0 and no row scalings");
  if (g_peers.mc && (!p.vec_ok || !aligned16(g_peers.mc)))
    return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "multicast target needs a 16-byte aligned W with even leading dimension");
  static dmma::PeerMaps pmaps_unused;   // the non-peer kernels ignore it
  dmma::PeerMaps& pmaps = pmaps_unused;
  // fused all-gather with bulk stores: one 3-D {M, N, batch} map per destination, box = one output tile
  dmma::PeerMaps peer_maps;
  const bool peer_bulk = g_peers.n > 0 && nscale == 0 && beta == 0.0 && p.vec_ok && !transA;
  const bool mc_staged = g_peers.mc != nullptr && !transA;
  if (peer_bulk || mc_staged) {
    for (int q = 0; peer_bulk && q <= g_peers.n; q++) {
      double* base = q == 0 ? C : static_cast<double*>(g_peers.ptr[q - 1]);
      cuuint64_t gdim[3] = {(cuuint64_t)M, (cuuint64_t)N, (cuuint64_t)batch};
      cuuint64_t gstr[2] = {(cuuint64_t)ldc * 8, batch > 1 ? (cuuint64_t)sC * 8 : (cuuint64_t)ldc * 8 * (cuuint64_t)N};
      cuuint32_t box[3] = {dmma::BM, dmma::BN, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = c->encode_tiled(&peer_maps.m[q], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, gdim, gstr, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return scimlop_fail(c, SCIMLOP_ERR_CUDA, "cuTensorMapEncodeTiled(peer C %d) failed (CUresult %d)", q, (int)r);
    }
    if (!peer_bulk) memset(&peer_maps, 0, sizeof(peer_maps));
    int grid = (int)std::min<long long>(p.total_tiles, std::max(1, c->sm_count - c->sm_reserve));
    if (c->grid_limit > 0) grid = std::min(grid, c->grid_limit);
    if (!transB) dmma::gemm_dmma_kernel<false, false, false, true><<<grid, dmma::THREADS, dmma::PEER_SMEM_BYTES, st>>>(tmA, tmB, p, peer_maps);
    else dmma::gemm_dmma_kernel<false, true, false, true><<<grid, dmma::THREADS, dmma::PEER_SMEM_BYTES, st>>>(tmA, tmB, p, peer_maps);
    c->launches++;
    SCIMLOP_CUDA(c, cudaGetLastError());
    return SCIMLOP_OK;
  }
  // few tiles (vector-input Kronecker products, small operators): split K over a thread-block cluster.  Pick the
  // cluster size with the smallest modelled time: waves x (k-blocks per CTA x 2.2 us + reduction); the number of
  // co-resident clusters comes from the occupancy query at handle creation (8-CTA clusters do not tile all SMs).
  const long long nk_all = (K + dmma::BK - 1) / dmma::BK;
  long long ks = 1;
  if (!g_peers.active() && !c->no_small_splitk && p.total_tiles < c->sm_count) {
    double best = (double)((p.total_tiles + c->sm_count - 1) / c->sm_count) * (nk_all * 2.2 + 2.0);
    for (long long cand = 2; cand <= 8; cand *= 2) {
      const long long maxc = c->dmma_max_clusters[cand];
      if (maxc <= 0 || nk_all < 2 * cand) continue;
      const double t = (double)((p.total_tiles + maxc - 1) / maxc) * ((double)((nk_all + cand - 1) / cand) * 2.2 + 6.0);
      if (t < best) { best = t; ks = cand; }
    }
  }
  if (!g_peers.active() && ks >= 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(ks * p.total_tiles), 1, 1);
    cfg.blockDim = dim3(dmma::THREADS, 1, 1);
    cfg.dynamicSmemBytes = dmma::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)ks;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (!transA && !transB) e = cudaLaunchKernelEx(&cfg, dmma::gemm_dmma_kernel<false, false, true>, tmA, tmB, p, pmaps);
    else if (!transA && transB) e = cudaLaunchKernelEx(&cfg, dmma::gemm_dmma_kernel<false, true, true>, tmA, tmB, p, pmaps);
    else if (transA && !transB) e = cudaLaunchKernelEx(&cfg, dmma::gemm_dmma_kernel<true, false, true>, tmA, tmB, p, pmaps);
    else e = cudaLaunchKernelEx(&cfg, dmma::gemm_dmma_kernel<true, true, true>, tmA, tmB, p, pmaps);
    c->launches++;
    SCIMLOP_CUDA(c, e);
    return SCIMLOP_OK;
  }
  // Tail-wave quantisation: a few whole waves plus a partial one (512 tiles on 148 SMs = 3.46 waves: the strong-scaled
  // config-2 share of one of 8 GPUs) leaves most SMs idle for the last tile time.  The last partial wave is then run as
  // a second launch of the split-K form -- each tile on a cluster of `ts` CTAs that divide K -- so it costs 1 / ts of a
  // tile time (+ the cluster reduction).  Modelled with the same constants as above; only taken when it saves > 10 us.
  int nsm = std::max(1, c->sm_count - c->sm_reserve);
  if (c->grid_limit > 0) nsm = std::min(nsm, c->grid_limit);
  long long tail = 0, ts = 1;
  if (c->tail_split && !g_peers.active() && p.total_tiles > nsm && p.total_tiles % nsm != 0) {
    const long long rem = p.total_tiles % nsm;
    const double tile_us = nk_all * 2.2 + 2.0;
    double best_gain = 10.0;
    for (long long cand = 2; cand <= 8; cand *= 2) {
      const long long maxc = c->dmma_max_clusters[cand];
      if (maxc <= 0 || nk_all < 2 * cand || rem > maxc) continue;
      const double gain = tile_us - ((double)((nk_all + cand - 1) / cand) * 2.2 + 6.0 + 3.0);
      if (gain > best_gain) { best_gain = gain; tail = rem; ts = cand; }
    }
  }
  const long long all_tiles = p.total_tiles;
  p.total_tiles = all_tiles - tail;
  const int grid = (int)std::min<long long>(p.total_tiles, nsm);
  if (!transA && !transB) dmma::gemm_dmma_kernel<false, false><<<grid, dmma::THREADS, dmma::SMEM_BYTES, st>>>(tmA, tmB, p, pmaps);
  else if (!transA && transB) dmma::gemm_dmma_kernel<false, true><<<grid, dmma::THREADS, dmma::SMEM_BYTES, st>>>(tmA, tmB, p, pmaps);
  else if (transA && !transB) dmma::gemm_dmma_kernel<true, false><<<grid, dmma::THREADS, dmma::SMEM_BYTES, st>>>(tmA, tmB, p, pmaps);
  else dmma::gemm_dmma_kernel<true, true><<<grid, dmma::THREADS, dmma::SMEM_BYTES, st>>>(tmA, tmB, p, pmaps);
  c->launches++;
  SCIMLOP_CUDA(c, cudaGetLastError());
  if (tail > 0) {
    p.tile_base = (int)(all_tiles - tail);
    p.total_tiles = tail;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(ts * tail), 1, 1);
    cfg.blockDim = dim3(dmma::THREADS, 1, 1);
    cfg.dynamicSmemBytes = dmma::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)ts;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (!transA && !transB) e = cudaLaunchKernelEx(&cfg, dmma::gemm_dmma_kernel<false, false, true>, tmA, tmB, p, pmaps);
    else if (!transA && transB) e = cudaLaunchKernelEx(&cfg, dmma::gemm_dmma_kernel<false, true, true>, tmA, tmB, p, pmaps);
    else if (transA && !transB) e = cudaLaunchKernelEx(&cfg, dmma::gemm_dmma_kernel<true, false, true>, tmA, tmB, p, pmaps);
    else e = cudaLaunchKernelEx(&cfg, dmma::gemm_dmma_kernel<true, true, true>, tmA, tmB, p, pmaps);
    c->launches++;
    SCIMLOP_CUDA(c, e);
  }
  return SCIMLOP_OK;
}

template <typename T>
int launch_simt(scimlop_context* c, bool transA, bool transB, long long M, long long N, long long K, T alpha,
                const T* A, long long lda, long long sA, const T* B, long long ldb, long long sB, T beta, T* C,
                long long ldc, long long sC, long long batch, int nscale, const ScaleArg* scales, cudaStream_t st) {
  simt::Params<T> p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.batch = (int)batch;
  p.tiles_m = (int)((M + simt::TM - 1) / simt::TM);
  p.tiles_n = (int)((N + simt::TN - 1) / simt::TN);
  const long long total = (long long)p.tiles_m * p.tiles_n * batch;
  if (total > 2147483647LL) return scimlop_fail(c, SCIMLOP_ERR_BAD_SHAPE, "gemm: too many tiles (%lld)", total);
  p.A = A; p.lda = lda; p.strideA = sA;
  p.B = B; p.ldb = ldb; p.strideB = sB;
  p.C = C; p.ldc = ldc; p.strideC = sC;
  p.alpha = alpha; p.beta = beta; p.nscale = nscale;
  for (int q = 0; q < nscale; q++) { p.scale[q].ptr = static_cast<const T*>(scales[q].ptr); p.scale[q].ld = scales[q].ld; }
  const unsigned grid = (unsigned)total;
  if (!transA && !transB) simt::gemm_simt_kernel<T, false, false><<<grid, simt::THREADS, 0, st>>>(p);
  else if (!transA && transB) simt::gemm_simt_kernel<T, false, true><<<grid, simt::THREADS, 0, st>>>(p);
  else if (transA && !transB) simt::gemm_simt_kernel<T, true, false><<<grid, simt::THREADS, 0, st>>>(p);
  else simt::gemm_simt_kernel<T, true, true><<<grid, simt::THREADS, 0, st>>>(p);
  c->launches++;
  SCIMLOP_CUDA(c, cudaGetLastError());
  return SCIMLOP_OK;
}

// ---- few-column products (k < 16): TMA-streamed, cluster-split, HBM-bound (gemm_skinny.cuh) ------------------
template <typename T>
bool skinny_eligible(scimlop_context* c, bool transB, long long M, long long N, long long K, const T* A, long long lda,
                     const T* B, long long ldb, long long batch, int nscale) {
  // up to 48 (Float64) / 64 (Float32) columns in passes of 8: measured cheaper than a 128-wide GEMM tile that is
  // mostly padding (16384^2 Float64: 0.30 ms per pass vs 2.3 ms for the DMMA kernel at any N <= 128)
  if (batch != 1 || transB || nscale != 0 || N < 1 || N > (sizeof(T) == 8 ? 48 : 64)) return false;
  if (M < 16 || K < 64 || M * K < (1LL << 16)) return false;           // tiny products: the generic path is fine
  const long long per16 = 16 / (long long)sizeof(T);
  if (!aligned16(A) || !aligned16(B) || lda % per16 || (N > 1 && ldb % per16)) return false;
  if (M > (1LL << 31) - 1 || K > (1LL << 31) - 1 || lda > (1LL << 36) || ldb > (1LL << 36)) return false;
  return c->encode_tiled != nullptr;
}

template <typename T, bool TRANS, int KC>
int launch_skinny_k(scimlop_context* c, const CUtensorMap& tmA, const CUtensorMap& tmX, const skinny::Params& p, dim3 grid,
                    cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(skinny::THREADS, 1, 1);
  cfg.dynamicSmemBytes = skinny::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = grid.x;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  auto kern = skinny::gemm_skinny_kernel<T, TRANS, KC>;
  SCIMLOP_CUDA(c, cudaLaunchKernelEx(&cfg, kern, tmA, tmX, p));
  c->launches++;
  return SCIMLOP_OK;
}

template <typename T>
int launch_skinny(scimlop_context* c, bool transA, long long M, long long N, long long K, T alpha, const T* A, long long lda,
                  const T* B, long long ldb, T beta, T* C, long long ldc, cudaStream_t st) {
  constexpr int VEC = 16 / (int)sizeof(T);
  constexpr int TC = 128 * VEC;
  const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const int kc = N == 1 ? 1 : (N == 2 ? 2 : (N <= 4 ? 4 : 8));
  CUtensorMap tmA, tmX;
  cuuint32_t estr[2] = {1, 1};
  {
    // stored matrix: contiguous extent first (NT: M rows; T: the reduction index)
    cuuint64_t gdim[2] = {(cuuint64_t)(transA ? K : M), (cuuint64_t)(transA ? M : K)};
    cuuint64_t gstr[1] = {(cuuint64_t)lda * sizeof(T)};
    cuuint32_t box[2] = {256, skinny::TO};
    CUresult r = c->encode_tiled(&tmA, dt, 2, const_cast<T*>(A), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return scimlop_fail(c, SCIMLOP_ERR_CUDA, "cuTensorMapEncodeTiled(skinny A) failed (CUresult %d)", (int)r);
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t gstr[1] = {(cuuint64_t)(N > 1 ? ldb : ((K * (long long)sizeof(T) + 15) / 16 * 16 / (long long)sizeof(T))) * sizeof(T)};
    cuuint32_t box[2] = {(cuuint32_t)(transA ? 256 : skinny::TO), (cuuint32_t)kc};
    CUresult r = c->encode_tiled(&tmX, dt, 2, const_cast<T*>(B), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return scimlop_fail(c, SCIMLOP_ERR_CUDA, "cuTensorMapEncodeTiled(skinny X) failed (CUresult %d)", (int)r);
  }
  skinny::Params p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = K; p.k = (int)N; p.ldy = ldc; p.Y = C; p.alpha = (double)alpha; p.beta = (double)beta;
  const long long otiles = transA ? (M + skinny::TO - 1) / skinny::TO : (M + TC - 1) / TC;
  p.tiles_red = (int)(transA ? (K + TC - 1) / TC : (K + skinny::TO - 1) / skinny::TO);
  const int passes = (int)((N + skinny::MAXC - 1) / skinny::MAXC);
  if (otiles > 65535) return -1;   // caller falls back
  long long s = c->sm_count / std::max<long long>(1, otiles * passes);
  s = std::min<long long>(s, std::max(1, p.tiles_red / 8));
  s = std::max<long long>(1, std::min<long long>(8, s));
  p.splits = (int)s;
  dim3 grid((unsigned)s, (unsigned)otiles, (unsigned)passes);
#define SCIMLOP_SKINNY(KC_)                                                                         \
  (transA ? launch_skinny_k<T, true, KC_>(c, tmA, tmX, p, grid, st) : launch_skinny_k<T, false, KC_>(c, tmA, tmX, p, grid, st))
  switch (kc) {
    case 1: return SCIMLOP_SKINNY(1);
    case 2: return SCIMLOP_SKINNY(2);
    case 4: return SCIMLOP_SKINNY(4);
    default: return SCIMLOP_SKINNY(8);
  }
#undef SCIMLOP_SKINNY
}

// C_b = alpha * rowscale .* (op(A_b) op(B_b)) + beta * C_b
template <typename T>
int gemm_dispatch(scimlop_context* c, int precision, bool transA, bool transB, long long M, long long N,
                  long long K, T alpha, const T* A, long long lda, long long sA, const T* B, long long ldb,
                  long long sB, T beta, T* C, long long ldc, long long sC, long long batch, int nscale,
                  const ScaleArg* scales, cudaStream_t st, const PeerArg& peers = NO_PEERS);

template <>
int gemm_dispatch<double>(scimlop_context* c, int precision, bool transA, bool transB, long long M, long long N,
                          long long K, double alpha, const double* A, long long lda, long long sA,
                          const double* B, long long ldb, long long sB, double beta, double* C, long long ldc,
                          long long sC, long long batch, int nscale, const ScaleArg* scales, cudaStream_t st,
                          const PeerArg& g_peers) {
  if (precision != SCIMLOP_PREC_DEFAULT)
    return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "Float64 only supports SCIMLOP_PREC_DEFAULT (got %d)", precision);
  if (M <= 0 || N <= 0 || batch <= 0) return SCIMLOP_OK;
  if (!g_peers.active() && skinny_eligible<double>(c, transB, M, N, K, A, lda, B, ldb, batch, nscale)) {
    const int rc = launch_skinny<double>(c, transA, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, st);
    if (rc >= 0) return rc;
  }
  if (g_peers.active() && !dmma_eligible(M, N, K, A, lda, sA, B, ldb, sB, batch))
    return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "peer targets need a DMMA-eligible (aligned Float64) last stage");
  if (dmma_eligible(M, N, K, A, lda, sA, B, ldb, sB, batch))
    return launch_dmma(c, transA, transB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, nscale,
                       scales, st, g_peers);
  return launch_simt<double>(c, transA, transB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch,
                             nscale, scales, st);
}

// ---- tcgen05 TF32 launcher (Float32, streamed operand x resident factor) ----------------------------------
int make_map_f32(scimlop_context* c, CUtensorMap* map, const float* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2,
                 cuuint64_t s1_bytes, cuuint64_t s2_bytes, cuuint32_t b0, cuuint32_t b1, bool mn_major,
                 bool swizzle = true) {
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstr[2] = {s1_bytes, d2 > 1 ? s2_bytes : s1_bytes * d1};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = c->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE,
                               // 32-bit MN-major UMMA operands need the 32-byte-atom flavour of the 128B swizzle
                               !swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE
                                        : (mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return scimlop_fail(c, SCIMLOP_ERR_CUDA, "cuTensorMapEncodeTiled(f32) failed (CUresult %d): dims {%llu,%llu,%llu}",
                        (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
  return SCIMLOP_OK;
}

// D(r,c,b) = alpha * sum_k X(r,k,b) * Fop(c,k) + beta * D;  x_mn: X(r,k) at r + ldx*k, else at k + ldx*r;
// f_k: Fop(c,k) at k + ldf*c, else at c + ldf*k.
int launch_tf32(scimlop_context* c, bool x_mn, bool f_k, long long rows, int m, int n, long long batch, const float* X,
                long long ldx, long long xbs, const float* F, long long ldf, float* D, long long d_rs, long long d_cs,
                long long d_bs, float alpha, float beta, int passes, cudaStream_t st) {
  CUtensorMap tmX, tmF, tmD;
  const bool xb = batch > 1 && xbs != 0;
  int rc;
  if (x_mn) rc = make_map_f32(c, &tmX, X, rows, n, xb ? batch : 1, ldx * 4, xbs * 4, 32, 32, true);
  else rc = make_map_f32(c, &tmX, X, n, rows, xb ? batch : 1, ldx * 4, xbs * 4, 32, 128, false);
  if (rc) return rc;
  if (f_k) rc = make_map_f32(c, &tmF, F, n, m, 1, ldf * 4, 0, 32, 128, false);
  else rc = make_map_f32(c, &tmF, F, m, n, 1, ldf * 4, 0, 32, 128, true);
  if (rc) return rc;
  tf32::Params p;
  memset(&p, 0, sizeof(p));
  p.rows = rows; p.m = m; p.n = n; p.batch = (int)batch;
  p.tiles_r = (int)((rows + tf32::TM - 1) / tf32::TM);
  p.tiles_c = 1;
  p.total_tiles = (long long)p.tiles_r * batch;
  p.D = D; p.d_rs = d_rs; p.d_cs = d_cs; p.d_bs = d_bs;
  p.alpha = alpha; p.beta = beta;
  p.vec_ok = (d_cs == 1) && aligned16(D) && (d_rs % 4 == 0) && (d_bs % 4 == 0);
  p.x_batched = xb;
  p.passes = passes;
  // epilogue through shared-memory staging + bulk tensor loads/stores when D is TMA-addressable
  p.tma_store = aligned16(D) && (x_mn ? (d_cs % 4 == 0 && (batch == 1 || d_bs % 4 == 0)) : (d_rs % 4 == 0));
  if (x_mn) rc = make_map_f32(c, &tmD, D, rows, m, batch, d_cs * 4, d_bs * 4, 32, 32, false, /*swizzle=*/false);
  else rc = make_map_f32(c, &tmD, D, m, rows, 1, d_rs * 4, 0, 32, 32, false, /*swizzle=*/true);
  if (!p.tma_store) tmD = tmX;   // unused by the kernel
  else if (rc) return rc;
  int grid = (int)std::min<long long>(p.total_tiles, std::max(1, c->sm_count - c->sm_reserve));
  if (c->grid_limit > 0) grid = std::min(grid, c->grid_limit);
  if (x_mn && f_k) tf32::gemm_tf32x3_kernel<true, true><<<grid, tf32::THREADS, tf32::SMEM_BYTES, st>>>(tmX, tmF, tmF, tmD, p);
  else if (x_mn) tf32::gemm_tf32x3_kernel<true, false><<<grid, tf32::THREADS, tf32::SMEM_BYTES, st>>>(tmX, tmF, tmF, tmD, p);
  else if (f_k) tf32::gemm_tf32x3_kernel<false, true><<<grid, tf32::THREADS, tf32::SMEM_BYTES, st>>>(tmX, tmF, tmF, tmD, p);
  else tf32::gemm_tf32x3_kernel<false, false><<<grid, tf32::THREADS, tf32::SMEM_BYTES, st>>>(tmX, tmF, tmF, tmD, p);
  c->launches++;
  SCIMLOP_CUDA(c, cudaGetLastError());
  return SCIMLOP_OK;
}

// ---- two fused Kronecker mode products, Float32 (kron2_f16.cuh) -----------------------------------------------------
// out_s(m1 x m2) = alpha * F1op(m1 x n1) * X_s(n1 x n2) * F2op(m2 x n2)^T + beta * out_s for `slabs` contiguous slabs.
// tr1 / tr2: the factor is the lazy transpose of the stored matrix (stored n x m).
#ifdef SCIMLOP_KF16_TRACE
long long* g_kf16_trace = nullptr;   // profiling builds only (tools/kf16_trace.py): 64 x 64 device stamps of CTA 0
#endif
bool kron2_f16_eligible(int precision, long long m1, long long n1, long long m2, long long n2, long long slabs,
                        const float* X, const float* D) {
  if (precision != SCIMLOP_PREC_DEFAULT && precision != SCIMLOP_PREC_TF32X3 && precision != SCIMLOP_PREC_BF16) return false;
  if (m1 > 128 || n1 > 128 || m2 > 128 || n2 > 128) return false;
  // every slab costs two 128^3 products whatever its extents: below ~96^2 the one-mode-per-pass path is faster
  if (n1 * n2 < 96 * 96 || m1 * m2 < 96 * 96) return false;
  if (n1 % 4 || m1 % 4 || !aligned16(X) || !aligned16(D)) return false;   // TMA strides are multiples of 16 bytes
  return slabs > 0 && slabs < (1ll << 31);
}

int launch_kron2_f16(scimlop_context* c, int m1, int n1, int m2, int n2, long long slabs, const float* F1, long long ld1,
                     bool tr1, const float* F2, long long ld2, bool tr2, const float* X, float* D, float alpha, float beta,
                     bool bf16, cudaStream_t st) {
  CUtensorMap tmX, tmW;
  if (int rc = make_map_f32(c, &tmX, X, n1, n2, slabs, (cuuint64_t)n1 * 4, (cuuint64_t)n1 * n2 * 4, 32, 128, false)) return rc;
  // beta != 0: the old output is prefetched as 128 x 32 blocks (plain layout: the epilogue lanes read consecutive floats)
  if (int rc = make_map_f32(c, &tmW, D, m1, m2, slabs, (cuuint64_t)m1 * 4, (cuuint64_t)m1 * m2 * 4, 128, 32, false, false)) return rc;
  kf16::Params p;
  memset(&p, 0, sizeof(p));
  p.n1 = n1; p.n2 = n2; p.m1 = m1; p.m2 = m2; p.slabs = slabs;
  p.F1 = F1; p.f1_rs = tr1 ? ld1 : 1; p.f1_cs = tr1 ? 1 : ld1;
  p.F2 = F2; p.f2_rs = tr2 ? ld2 : 1; p.f2_cs = tr2 ? 1 : ld2;
  p.D = D; p.alpha = alpha; p.beta = beta;
#ifdef SCIMLOP_KF16_TRACE
  p.trace = g_kf16_trace;
#endif
  int grid = (int)std::min<long long>(slabs, std::max(1, c->sm_count - c->sm_reserve));
  if (c->grid_limit > 0) grid = std::min(grid, c->grid_limit);
  if (bf16 && beta != 0.f) kf16::kron2_f16_kernel<true, true><<<grid, kf16::THREADS, kf16::SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (bf16) kf16::kron2_f16_kernel<false, true><<<grid, kf16::THREADS, kf16::SMEM_BYTES, st>>>(tmX, tmW, p);
  else if (beta != 0.f) kf16::kron2_f16_kernel<true><<<grid, kf16::THREADS, kf16::SMEM_BYTES, st>>>(tmX, tmW, p);
  else kf16::kron2_f16_kernel<false><<<grid, kf16::THREADS, kf16::SMEM_BYTES, st>>>(tmX, tmW, p);
  c->launches++;
  SCIMLOP_CUDA(c, cudaGetLastError());
  return SCIMLOP_OK;
}

// ---- two fused Kronecker mode products, Float64, small problems (kron2_f64_small.cuh) -----------------------------------
// Taken when the slab count is too small to fill the persistent 128 x 128-tile GEMM (fewer than two waves of tiles):
// config 1, vector-like batches of small-factor products.
bool kron2_f64_small_eligible(const scimlop_context* c, long long m1, long long n1, long long m2, long long n2, long long slabs) {
  if (m1 > k2s::MAXD || n1 > k2s::MAXD || m2 > k2s::MAXD || n2 > k2s::MAXD) return false;
  if (m1 < 8 || n1 < 8 || m2 < 8 || n2 < 8) return false;
  return slabs >= 1 && slabs <= 2ll * c->sm_count;
}
int launch_kron2_f64_small(scimlop_context* c, int m1, int n1, int m2, int n2, long long slabs, const double* F1, long long ld1,
                           bool tr1, const double* F2, long long ld2, bool tr2, const double* X, double* D, double alpha,
                           double beta, cudaStream_t st) {
  k2s::Params p;
  memset(&p, 0, sizeof(p));
  p.n1 = n1; p.n2 = n2; p.m1 = m1; p.m2 = m2; p.slabs = slabs;
  p.F1 = F1; p.f1_rs = tr1 ? ld1 : 1; p.f1_cs = tr1 ? 1 : ld1;
  p.F2 = F2; p.f2_rs = tr2 ? ld2 : 1; p.f2_cs = tr2 ? 1 : ld2;
  p.V = X; p.W = D; p.alpha = alpha; p.beta = beta;
  const size_t smem = k2s::smem_bytes(n1, n2, m1, m2);
  const int grid = (int)std::min<long long>(slabs, c->sm_count);
  k2s::kron2_f64_small_kernel<<<grid, k2s::THREADS, smem, st>>>(p);
  c->launches++;
  SCIMLOP_CUDA(c, cudaGetLastError());
  return SCIMLOP_OK;
}

// ---- hi / lo images of Float32 operands (3xTF32 split) ---------------------------------------------------------
inline long long split_ld(long long rows) { return (rows + 3) / 4 * 4; }
inline size_t split_image_bytes(long long rows, long long cols) { return 2 * (size_t)split_ld(rows) * (size_t)cols * sizeof(float); }

int launch_split(scimlop_context* c, const float* src, long long ld, long long rows, long long cols, float* hi, float* lo,
                 cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return SCIMLOP_OK;
  const bool vec = aligned16(src) && ld % 4 == 0;
  const long long items = ((vec ? rows / 4 : 0) + (rows - (vec ? rows / 4 * 4 : 0))) * cols;
  const int grid = (int)std::max<long long>(1, std::min<long long>((items + 255) / 256, (long long)c->sm_count * 8));
  if (vec) tf32::split_hilo_kernel<true><<<grid, 256, 0, st>>>(src, ld, rows, cols, hi, lo, split_ld(rows));
  else tf32::split_hilo_kernel<false><<<grid, 256, 0, st>>>(src, ld, rows, cols, hi, lo, split_ld(rows));
  c->launches++;
  SCIMLOP_CUDA(c, cudaGetLastError());
  return SCIMLOP_OK;
}

const scimlop_split_entry* find_split(const scimlop_context* c, const void* src, long long ld, long long rows, long long cols) {
  for (int i = 0; i < c->nsplits; i++) {
    const scimlop_split_entry& e = c->splits[i];
    if (e.src == src && e.ld == ld && e.rows == rows && e.cols == cols) return &e;
  }
  return nullptr;
}

// General Float32 GEMM on tcgen05 (3xTF32): C_b(M x N) = alpha * op(A_b)(M x K) * op(B)(K x N) + beta * C_b, any K, N.
// One operand ("X") streams through the splitter warps into TMEM, the other ("F") streams through the same ring as
// 128-column tiles of its hi / lo halves, which the CALLER prepared once (scimlop_split_attach at cache_operator
// time, refreshed by update_coefficients!; `pre`), or -- for the uncached convenience entries -- which are split here
// into the handle's grow-only scratch (one extra launch; the first call of a shape allocates).
// `f_is_a == false`: X = op(A), F = op(B), rows = M (output contiguous along r; batches of A / C with one shared B
// allowed); `f_is_a == true`: X = op(B)^T, F = op(A), rows = N (output contiguous along c).
int launch_tf32_stream(scimlop_context* c, bool f_is_a, const scimlop_split_entry* pre, bool transA, bool transB,
                       long long M, long long N, long long K, float alpha, const float* A, long long lda, long long sA,
                       const float* B, long long ldb, float beta, float* C, long long ldc, long long sC, long long batch,
                       int passes, int nscale, const ScaleArg* scales, cudaStream_t st) {
  const float* Fsrc = f_is_a ? A : B;
  const long long ldsrc = f_is_a ? lda : ldb;
  const long long frows = f_is_a ? (transA ? K : M) : (transB ? N : K);   // stored rows of F
  const long long fcols = f_is_a ? (transA ? M : K) : (transB ? K : N);   // stored columns of F
  const float *Fhi, *Flo;
  long long ldf;
  if (pre) {
    Fhi = pre->hi; Flo = pre->lo; ldf = pre->ld_split;
  } else {
    const size_t need = split_image_bytes(frows, fcols);
    if (need > c->split_scratch_bytes) {
      // uncached entry: grow the handle-owned scratch (synchronises; cached operators never come here)
      SCIMLOP_CUDA(c, cudaDeviceSynchronize());
      if (c->split_scratch) { SCIMLOP_CUDA(c, cudaFree(c->split_scratch)); c->split_scratch = nullptr; c->split_scratch_bytes = 0; }
      SCIMLOP_CUDA(c, cudaMalloc(&c->split_scratch, need));
      c->split_scratch_bytes = need;
    }
    ldf = split_ld(frows);
    float* hi = static_cast<float*>(c->split_scratch);
    float* lo = hi + (size_t)ldf * (size_t)fcols;
    if (int rc = launch_split(c, Fsrc, ldsrc, frows, fcols, hi, lo, st)) return rc;
    Fhi = hi; Flo = lo;
  }
  const size_t half = (size_t)ldf * (size_t)fcols * sizeof(float);
  // roles
  const long long rows = f_is_a ? N : M, mcols = f_is_a ? M : N;
  const float* X = f_is_a ? B : A;
  const long long ldx = f_is_a ? ldb : lda;
  // X(r, k): !f_is_a: A(m, k): transA ? K-major : MN-major;  f_is_a: B(k, n): transB (stored N x K) ? MN-major : K-major
  const bool x_mn = f_is_a ? transB : !transA;
  // F(c, k): !f_is_a: B(k, n): transB ? MN-major : K-major;   f_is_a: A(m, k): transA (stored K x M) ? K-major : MN-major
  const bool f_k = f_is_a ? transA : !transB;
  const bool xb = !f_is_a && batch > 1 && sA != 0;
  CUtensorMap tmX, tmFhi, tmFlo, tmD;
  int rc;
  if (x_mn) rc = make_map_f32(c, &tmX, X, rows, K, xb ? batch : 1, ldx * 4, sA * 4, 32, 32, true);
  else rc = make_map_f32(c, &tmX, X, K, rows, xb ? batch : 1, ldx * 4, sA * 4, 32, 128, false);
  for (int h = 0; h < 2 && !rc; h++) {
    const float* base = h ? Flo : Fhi;
    CUtensorMap* tm = h ? &tmFlo : &tmFhi;
    if (f_k) rc = make_map_f32(c, tm, base, K, mcols, 1, ldf * 4, 0, 32, 128, false);
    else rc = make_map_f32(c, tm, base, mcols, K, 1, ldf * 4, 0, 32, 32, true);
  }
  if (rc) return rc;
  tf32::Params p;
  memset(&p, 0, sizeof(p));
  p.rows = rows; p.m = (int)mcols; p.n = (int)K; p.batch = (int)batch;
  p.tiles_r = (int)((rows + tf32::TM - 1) / tf32::TM);
  p.tiles_c = (int)((mcols + 127) / 128);
  p.total_tiles = (long long)p.tiles_r * p.tiles_c * batch;
  // a small pre-split operand (a Kronecker factor) stays in L2: then each X row tile is fetched once.  (Measured:
  // with a 64 MB pre-split V the column-fastest order is 40 % slower than row-fastest, so the bound is tight.)
  p.c_fast = 2 * half <= (16u << 20);
  p.D = C;
  p.d_rs = f_is_a ? ldc : 1; p.d_cs = f_is_a ? 1 : ldc; p.d_bs = sC;
  p.alpha = alpha; p.beta = beta;
  p.vec_ok = f_is_a && aligned16(C) && ldc % 4 == 0;
  p.x_batched = xb;
  p.passes = passes;
  p.nscale = nscale; p.rows_are_n = f_is_a;
  for (int q = 0; q < nscale; q++) { p.scale_ptr[q] = static_cast<const float*>(scales[q].ptr); p.scale_ld[q] = scales[q].ld; }
  p.tma_store = aligned16(C) && ldc % 4 == 0 && (batch == 1 || sC % 4 == 0);
  if (!p.tma_store) tmD = tmX;
  else if (!f_is_a) rc = make_map_f32(c, &tmD, C, M, N, batch, ldc * 4, sC * 4, 32, 32, false, /*swizzle=*/false);
  else rc = make_map_f32(c, &tmD, C, M, N, 1, ldc * 4, 0, 32, 32, false, /*swizzle=*/true);
  if (rc) return rc;
  const int grid = (int)std::min<long long>(p.total_tiles, std::max(1, c->sm_count - c->sm_reserve));
#define SCIMLOP_TF32S(XM, FK, OR) \
  tf32::gemm_tf32x3_kernel<XM, FK, true, OR><<<grid, tf32::THREADS, tf32::SMEM_BYTES, st>>>(tmX, tmFhi, tmFlo, tmD, p)
  if (!f_is_a) {
    if (x_mn && f_k) SCIMLOP_TF32S(true, true, true);
    else if (x_mn) SCIMLOP_TF32S(true, false, true);
    else if (f_k) SCIMLOP_TF32S(false, true, true);
    else SCIMLOP_TF32S(false, false, true);
  } else {
    if (x_mn && f_k) SCIMLOP_TF32S(true, true, false);
    else if (x_mn) SCIMLOP_TF32S(true, false, false);
    else if (f_k) SCIMLOP_TF32S(false, true, false);
    else SCIMLOP_TF32S(false, false, false);
  }
#undef SCIMLOP_TF32S
  c->launches++;
  SCIMLOP_CUDA(c, cudaGetLastError());
  return SCIMLOP_OK;
}

template <>
int gemm_dispatch<float>(scimlop_context* c, int precision, bool transA, bool transB, long long M, long long N,
                         long long K, float alpha, const float* A, long long lda, long long sA, const float* B,
                         long long ldb, long long sB, float beta, float* C, long long ldc, long long sC,
                         long long batch, int nscale, const ScaleArg* scales, cudaStream_t st, const PeerArg& peers) {
  if (peers.active()) return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "peer / multicast targets are Float64 only");
  if (precision < SCIMLOP_PREC_DEFAULT || precision > SCIMLOP_PREC_BF16)
    return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "unknown precision mode %d", precision);
  if (precision == SCIMLOP_PREC_BF16) precision = SCIMLOP_PREC_TF32;   // outside the fused Kronecker kernel: single-pass TF32
  if (M <= 0 || N <= 0 || batch <= 0) return SCIMLOP_OK;
  if (precision != SCIMLOP_PREC_FP32_EXACT && nscale == 0 && K >= 8 && K <= tf32::MAXF) {
    const int passes = (precision == SCIMLOP_PREC_TF32) ? 1 : 3;
    const long long lim = 2147483647LL;
    // (a) the factor is A (innermost Kronecker mode, MatrixOperator mul!): stream B's columns
    if (batch == 1 && !transB && M <= tf32::MAXF && N >= 256 && N < lim && aligned16(A) && aligned16(B) &&
        lda % 4 == 0 && ldb % 4 == 0)
      return launch_tf32(c, /*x_mn=*/false, /*f_k=*/transA, N, (int)M, (int)K, 1, B, ldb, 0, A, lda, C, ldc, 1, 0, alpha,
                         beta, passes, st);
    // (b) the factor is B, shared by all batches (any other Kronecker mode): stream A's rows
    if (!transA && (sB == 0 || batch == 1) && N <= tf32::MAXF && M * batch >= 256 && M < lim && batch < lim &&
        aligned16(A) && aligned16(B) && lda % 4 == 0 && ldb % 4 == 0 && sA % 4 == 0)
      return launch_tf32(c, /*x_mn=*/true, /*f_k=*/!transB, M, (int)N, (int)K, batch, A, lda, sA, B, ldb, C, 1, ldc, sC,
                         alpha, beta, passes, st);
  }
  if (precision != SCIMLOP_PREC_TF32 && skinny_eligible<float>(c, transB, M, N, K, A, lda, B, ldb, batch, nscale)) {
    const int rc = launch_skinny<float>(c, transA, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, st);
    if (rc >= 0) return rc;
  }
  // large Float32 products: the streamed-operand tcgen05 kernel (the FFMA kernel below peaks at ~27 TFLOP/s).
  // The smaller operand is the one that gets pre-split; batches need one shared B (the outer Kronecker modes).
  if (precision != SCIMLOP_PREC_FP32_EXACT && (nscale == 0 || batch == 1) && K >= 32 && (batch == 1 || sB == 0) &&
      2.0 * (double)M * (double)N * (double)K * (double)batch >= 1e8 && M < 2147483647LL && N < 2147483647LL &&
      K < 2147483647LL && batch < 2147483647LL && aligned16(A) && aligned16(B) && lda % 4 == 0 && ldb % 4 == 0 &&
      (batch == 1 || sA % 4 == 0)) {
    const int passes = precision == SCIMLOP_PREC_TF32 ? 1 : 3;
    // an operand whose hi / lo image the caller attached (cached operators) is the pre-split one; otherwise the
    // smaller operand is split per call (uncached entries)
    const scimlop_split_entry* preA = batch == 1 ? find_split(c, A, lda, transA ? K : M, transA ? M : K) : nullptr;
    const scimlop_split_entry* preB = find_split(c, B, ldb, transB ? N : K, transB ? K : N);
    if (preA && !preB && N >= 16 && M >= 16)
      return launch_tf32_stream(c, true, preA, transA, transB, M, N, K, alpha, A, lda, 0, B, ldb, beta, C, ldc, 0, 1, passes,
                                nscale, scales, st);
    if (preB && M * batch >= 16 && N >= 16)
      return launch_tf32_stream(c, false, preB, transA, transB, M, N, K, alpha, A, lda, sA, B, ldb, beta, C, ldc, sC, batch,
                                passes, nscale, scales, st);
    const bool a_small = batch == 1 && (double)M * (double)K < (double)K * (double)N;
    if (a_small && N >= 128 && M >= 16)
      return launch_tf32_stream(c, true, preA, transA, transB, M, N, K, alpha, A, lda, 0, B, ldb, beta, C, ldc, 0, 1, passes,
                                nscale, scales, st);
    if (M * batch >= 128 && N >= 16)
      return launch_tf32_stream(c, false, nullptr, transA, transB, M, N, K, alpha, A, lda, sA, B, ldb, beta, C, ldc, sC, batch,
                                passes, nscale, scales, st);
  }
  if (precision == SCIMLOP_PREC_TF32)
    return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED,
                        "single-pass TF32 was requested but this shape is not eligible for the tcgen05 path");
  return launch_simt<float>(c, transA, transB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, nscale,
                            scales, st);
}

int check_handle(scimlop_context* c) { return scimlop_valid(c) ? SCIMLOP_OK : SCIMLOP_ERR_BAD_HANDLE; }
int check_dtype(scimlop_context* c, int dtype) {
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32)
    return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "unsupported dtype %d (Float64 = 0, Float32 = 1)", dtype);
  return SCIMLOP_OK;
}

}  // namespace

// ================================================================================================
// handle
// ================================================================================================
extern "C" int scimlop_version(void) { return SCIMLOP_VERSION; }
#ifdef SCIMLOP_KF16_TRACE
extern "C" int scimlop_debug_kf16_trace(void* buf) { g_kf16_trace = static_cast<long long*>(buf); return 0; }
#endif

extern "C" const char* scimlop_status_string(int s) {
  switch (s) {
    case SCIMLOP_OK: return "ok";
    case SCIMLOP_ERR_NULL_POINTER: return "null pointer";
    case SCIMLOP_ERR_BAD_SHAPE: return "bad shape";
    case SCIMLOP_ERR_WORKSPACE: return "workspace missing or too small (operator not cached)";
    case SCIMLOP_ERR_UNSUPPORTED: return "unsupported dtype / precision / factor kind";
    case SCIMLOP_ERR_CUDA: return "CUDA error";
    case SCIMLOP_ERR_NCCL: return "NCCL error";
    case SCIMLOP_ERR_BAD_HANDLE: return "bad handle";
    case SCIMLOP_ERR_NO_DEVICE: return "no sm_100 device (there is no CPU fallback)";
    case SCIMLOP_ERR_SINGULAR: return "singular matrix";
    default: return "unknown status";
  }
}

extern "C" int scimlop_create(scimlop_handle_t* out, int device) {
  if (!out) return SCIMLOP_ERR_NULL_POINTER;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return SCIMLOP_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SCIMLOP_ERR_NO_DEVICE;
  if (prop.major != 10) return SCIMLOP_ERR_NO_DEVICE;  // sm_100a cubins only
  scimlop_context* c = new (std::nothrow) scimlop_context();
  if (!c) return SCIMLOP_ERR_CUDA;
  memset(c, 0, sizeof(*c));
  c->magic = SCIMLOP_MAGIC;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->tail_split = 1;
  c->peer_pipeline = 1;
  scimlop_device_guard guard(device);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess) {
    delete c;
    return SCIMLOP_ERR_CUDA;
  }
  c->encode_tiled = reinterpret_cast<scimlop_encode_tiled_fn>(fn);
  cudaError_t e = cudaSuccess;
  auto optin = [&](const void* f) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, dmma::SMEM_BYTES);
  };
  optin((const void*)dmma::gemm_dmma_kernel<false, false>);
  optin((const void*)dmma::gemm_dmma_kernel<false, true>);
  optin((const void*)dmma::gemm_dmma_kernel<true, false>);
  optin((const void*)dmma::gemm_dmma_kernel<true, true>);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)dmma::gemm_dmma_kernel<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dmma::PEER_SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)dmma::gemm_dmma_kernel<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dmma::PEER_SMEM_BYTES);
  optin((const void*)dmma::gemm_dmma_kernel<false, false, true>);
  optin((const void*)dmma::gemm_dmma_kernel<false, true, true>);
  optin((const void*)dmma::gemm_dmma_kernel<true, false, true>);
  optin((const void*)dmma::gemm_dmma_kernel<true, true, true>);
  for (int cs = 2; cs <= 8 && e == cudaSuccess; cs *= 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(cs * 64, 1, 1);
    cfg.blockDim = dim3(dmma::THREADS, 1, 1);
    cfg.dynamicSmemBytes = dmma::SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, dmma::gemm_dmma_kernel<false, false, true>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      nclusters = 0;
    }
    c->dmma_max_clusters[cs] = nclusters;
  }
  auto optin2 = [&](const void* f) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, tf32::SMEM_BYTES);
  };
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)k2s::kron2_f64_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)k2s::smem_bytes(k2s::MAXD, k2s::MAXD, k2s::MAXD, k2s::MAXD));
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)kf16::kron2_f16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kf16::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)kf16::kron2_f16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kf16::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)kf16::kron2_f16_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kf16::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)kf16::kron2_f16_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kf16::SMEM_BYTES);
  optin2((const void*)tf32::gemm_tf32x3_kernel<false, false>);
  optin2((const void*)tf32::gemm_tf32x3_kernel<false, true>);
  optin2((const void*)tf32::gemm_tf32x3_kernel<true, false>);
  optin2((const void*)tf32::gemm_tf32x3_kernel<true, true, true, true>);
  optin2((const void*)tf32::gemm_tf32x3_kernel<true, false, true, true>);
  optin2((const void*)tf32::gemm_tf32x3_kernel<false, true, true, true>);
  optin2((const void*)tf32::gemm_tf32x3_kernel<false, false, true, true>);
  optin2((const void*)tf32::gemm_tf32x3_kernel<true, true, true, false>);
  optin2((const void*)tf32::gemm_tf32x3_kernel<true, false, true, false>);
  optin2((const void*)tf32::gemm_tf32x3_kernel<false, true, true, false>);
  optin2((const void*)tf32::gemm_tf32x3_kernel<false, false, true, false>);
  optin2((const void*)tf32::gemm_tf32x3_kernel<true, true>);
  auto optin3 = [&](const void* f) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, skinny::SMEM_BYTES);
  };
#define SCIMLOP_SKINNY_OPTIN(T_, KC_)                                      \
  optin3((const void*)skinny::gemm_skinny_kernel<T_, false, KC_>);        \
  optin3((const void*)skinny::gemm_skinny_kernel<T_, true, KC_>);
  SCIMLOP_SKINNY_OPTIN(double, 1) SCIMLOP_SKINNY_OPTIN(double, 2) SCIMLOP_SKINNY_OPTIN(double, 4) SCIMLOP_SKINNY_OPTIN(double, 8)
  SCIMLOP_SKINNY_OPTIN(float, 1) SCIMLOP_SKINNY_OPTIN(float, 2) SCIMLOP_SKINNY_OPTIN(float, 4) SCIMLOP_SKINNY_OPTIN(float, 8)
#undef SCIMLOP_SKINNY_OPTIN
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)band::band_kronsum_tri_tma_kernel<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, band::tri_smem_bytes<double>());
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)band::band_kronsum_tri_tma_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, band::tri_smem_bytes<double, true>());
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)band::band_kronsum_tri_tma_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, band::tri_smem_bytes<float>());
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)band::band_kronsum_tri_tma_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, band::tri_smem_bytes<float, true>());
  if (e != cudaSuccess) {
    delete c;
    return SCIMLOP_ERR_CUDA;
  }
  *out = c;
  return SCIMLOP_OK;
}

extern "C" int scimlop_last_error(scimlop_handle_t h, char* buf, size_t len) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (!buf || len == 0) return SCIMLOP_ERR_NULL_POINTER;
  snprintf(buf, len, "%s", h->last_error);
  return SCIMLOP_OK;
}

extern "C" int scimlop_set_option(scimlop_handle_t h, int option, int value) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (option == SCIMLOP_OPT_TAIL_SPLIT) { h->tail_split = value != 0; return SCIMLOP_OK; }
  if (option == SCIMLOP_OPT_MODE_PIPELINE) { h->mode_pipeline = (value >= 2 && value <= 8) ? value : 0; return SCIMLOP_OK; }
  if (option == SCIMLOP_OPT_PEER_CHUNKS) { h->peer_chunks = (value == 2 || value == 4 || value == 8) ? value : 0; return SCIMLOP_OK; }
  if (option == SCIMLOP_OPT_PEER_STAGE2_SMS) { h->peer_stage2_sms = value > 0 ? value : 0; return SCIMLOP_OK; }
  if (option == SCIMLOP_OPT_PEER_PIPELINE) { h->peer_pipeline = value < 0 ? 0 : (value > 2 ? 2 : value); return SCIMLOP_OK; }
  return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "scimlop_set_option: unknown option %d", option);
}

extern "C" int scimlop_launch_count(scimlop_handle_t h, int64_t* count) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (!count) return SCIMLOP_ERR_NULL_POINTER;
  *count = h->launches;
  return SCIMLOP_OK;
}

// ================================================================================================
// leaves
// ================================================================================================
extern "C" int scimlop_gemm_strided_batched(scimlop_handle_t h, int dtype, int precision, int transA, int transB,
                                            int64_t M, int64_t N, int64_t K, const void* alpha, const void* A,
                                            int64_t lda, int64_t strideA, const void* B, int64_t ldb,
                                            int64_t strideB, const void* beta, void* C, int64_t ldc,
                                            int64_t strideC, int64_t batch, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (int rc = check_dtype(h, dtype)) return rc;
  if (M < 0 || N < 0 || K < 0 || batch < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "gemm: negative size");
  if (M == 0 || N == 0 || batch == 0) return SCIMLOP_OK;
  if (!C || (K > 0 && (!A || !B))) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "gemm: null operand");
  const int64_t a_rows = transA ? K : M, b_rows = transB ? N : K;
  if ((K > 0 && (lda < std::max<int64_t>(1, a_rows) || ldb < std::max<int64_t>(1, b_rows))) || ldc < M)
    return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "gemm: leading dimension smaller than rows (lda=%lld ldb=%lld ldc=%lld)",
                        (long long)lda, (long long)ldb, (long long)ldc);
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SCIMLOP_F64)
    return gemm_dispatch<double>(h, precision, transA != 0, transB != 0, M, N, K, scalar_or<double>(alpha, 1.0),
                                 (const double*)A, lda, strideA, (const double*)B, ldb, strideB,
                                 scalar_or<double>(beta, 0.0), (double*)C, ldc, strideC, batch, 0, nullptr, st);
  return gemm_dispatch<float>(h, precision, transA != 0, transB != 0, M, N, K, scalar_or<float>(alpha, 1.f),
                              (const float*)A, lda, strideA, (const float*)B, ldb, strideB,
                              scalar_or<float>(beta, 0.f), (float*)C, ldc, strideC, batch, 0, nullptr, st);
}

extern "C" int scimlop_matmul(scimlop_handle_t h, int dtype, int precision, int transA, int64_t m, int64_t n,
                              int64_t k, const void* A, int64_t lda, const void* V, int64_t ldv, void* W,
                              int64_t ldw, const void* alpha, const void* beta, void* stream) {
  return scimlop_gemm_strided_batched(h, dtype, precision, transA, 0, m, k, n, alpha, A, lda, 0, V, ldv, 0, beta, W,
                                      ldw, 0, 1, stream);
}

// ---- hi / lo images of Float32 operands: prepared once per cached operator ---------------------------------------
extern "C" int scimlop_split_bytes(int64_t rows, int64_t cols, size_t* bytes) {
  if (!bytes) return SCIMLOP_ERR_NULL_POINTER;
  if (rows < 0 || cols < 0) return SCIMLOP_ERR_BAD_SHAPE;
  *bytes = split_image_bytes(rows, cols);
  return SCIMLOP_OK;
}

extern "C" int scimlop_split_attach(scimlop_handle_t h, const void* A, int64_t lda, int64_t rows, int64_t cols, void* split,
                                    size_t split_bytes, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!A || !split) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "split_attach: null pointer");
  if (rows <= 0 || cols <= 0 || lda < rows) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "split_attach: bad shape");
  if (split_bytes < split_image_bytes(rows, cols) || !aligned16(split))
    return scimlop_fail(h, SCIMLOP_ERR_WORKSPACE, "split_attach: image buffer too small (%zu < %zu B) or not 16-byte aligned",
                        split_bytes, split_image_bytes(rows, cols));
  scimlop_split_entry* e = nullptr;
  for (int i = 0; i < h->nsplits; i++)
    if (h->splits[i].src == A) e = &h->splits[i];     // re-attach replaces
  if (!e) {
    if (h->nsplits == SCIMLOP_MAX_SPLITS) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "split_attach: more than %d attached operands", SCIMLOP_MAX_SPLITS);
    e = &h->splits[h->nsplits++];
  }
  e->src = A; e->ld = lda; e->rows = rows; e->cols = cols;
  e->ld_split = split_ld(rows);
  e->hi = static_cast<float*>(split);
  e->lo = e->hi + (size_t)e->ld_split * (size_t)cols;
  scimlop_device_guard guard(h->device);
  return launch_split(h, static_cast<const float*>(A), lda, rows, cols, e->hi, e->lo, static_cast<cudaStream_t>(stream));
}

extern "C" int scimlop_split_refresh(scimlop_handle_t h, const void* A, void* stream) {
  if (int rc = check_handle(h)) return rc;
  for (int i = 0; i < h->nsplits; i++) {
    const scimlop_split_entry& e = h->splits[i];
    if (e.src != A) continue;
    scimlop_device_guard guard(h->device);
    return launch_split(h, static_cast<const float*>(A), e.ld, e.rows, e.cols, e.hi, e.lo, static_cast<cudaStream_t>(stream));
  }
  return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "split_refresh: operand %p is not attached", A);
}

extern "C" int scimlop_split_detach(scimlop_handle_t h, const void* A) {
  if (int rc = check_handle(h)) return rc;
  for (int i = 0; i < h->nsplits; i++)
    if (h->splits[i].src == A) {
      h->splits[i] = h->splits[--h->nsplits];
      return SCIMLOP_OK;
    }
  return SCIMLOP_OK;   // detaching an unknown operand is a no-op
}

template <typename T>
static int diag_mul_t(scimlop_context* h, int64_t n, int64_t k, const void* d, int64_t ldd, const void* V,
                      int64_t ldv, void* W, int64_t ldw, const void* alpha, const void* beta, cudaStream_t st) {
  elt::Params<T> p;
  memset(&p, 0, sizeof(p));
  p.n = n; p.k = k; p.V = (const T*)V; p.ldv = ldv; p.W = (T*)W; p.ldw = ldw;
  p.alpha = scalar_or<T>(alpha, T(1)); p.beta = scalar_or<T>(beta, T(0));
  p.nterms = 1; p.coef[0] = T(1); p.nfac[0] = 1;
  p.fac[0][0].ptr = (const T*)d;
  if (ldd == 0) { p.nvec[0] = 1; p.fac[0][0].ld = 0; p.fac[0][0].div = 1; p.fac[0][0].mod = n; }
  else { p.nvec[0] = 0; p.fac[0][0].ld = ldd; p.fac[0][0].div = 1; p.fac[0][0].mod = 1; }
  return launch_eltwise(h, p, st);
}

extern "C" int scimlop_diag_mul(scimlop_handle_t h, int dtype, int64_t n, int64_t k, const void* d, int64_t ldd,
                                const void* V, int64_t ldv, void* W, int64_t ldw, const void* alpha,
                                const void* beta, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (int rc = check_dtype(h, dtype)) return rc;
  if (n < 0 || k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "diag_mul: negative size");
  if (n == 0 || k == 0) return SCIMLOP_OK;
  if (!d || !V || !W) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "diag_mul: null operand");
  if (ldv < n || ldw < n || (ldd != 0 && ldd < n))
    return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "diag_mul: leading dimension smaller than n");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == SCIMLOP_F64 ? diag_mul_t<double>(h, n, k, d, ldd, V, ldv, W, ldw, alpha, beta, st)
                              : diag_mul_t<float>(h, n, k, d, ldd, V, ldv, W, ldw, alpha, beta, st);
}

extern "C" int scimlop_axpby(scimlop_handle_t h, int dtype, int64_t n, int64_t k, const void* V, int64_t ldv,
                             void* W, int64_t ldw, const void* alpha, const void* beta, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (int rc = check_dtype(h, dtype)) return rc;
  if (n < 0 || k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "axpby: negative size");
  if (n == 0 || k == 0) return SCIMLOP_OK;
  if (!W) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "axpby: W is null");
  if (ldw < n || (V && ldv < n)) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "axpby: leading dimension smaller than n");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SCIMLOP_F64)
    return eltwise_axpby<double>(h, n, k, (const double*)V, ldv, (double*)W, ldw, scalar_or<double>(alpha, 1.0),
                                 scalar_or<double>(beta, 0.0), st);
  return eltwise_axpby<float>(h, n, k, (const float*)V, ldv, (float*)W, ldw, scalar_or<float>(alpha, 1.f),
                              scalar_or<float>(beta, 0.f), st);
}

// ================================================================================================
// BlockDiagonalOperator (src/block.jl:145-179)
// ================================================================================================
namespace {

template <typename T>
int blockdiag_t(scimlop_context* h, int precision, int nblocks, const scimlop_factor_t* blk, const T* V, int64_t ldv,
                T* W, int64_t ldw, int64_t k, T alpha, T beta, cudaStream_t st) {
  // equal dense blocks at a uniform stride: ONE strided-batched GEMM, block b reads rows [b*n, (b+1)*n) of V and
  // writes rows [b*m, (b+1)*m) of W (row offsets are just batch strides of the column-major views)
  bool uniform = nblocks > 1 && (blk[0].kind & 0xFF) == SCIMLOP_DENSE;
  long long stride = 0;
  for (int b = 1; uniform && b < nblocks; b++) {
    uniform = blk[b].kind == blk[0].kind && blk[b].trans == blk[0].trans && blk[b].rows == blk[0].rows &&
              blk[b].cols == blk[0].cols && blk[b].ld == blk[0].ld;
    const long long d = static_cast<const char*>(blk[b].data) - static_cast<const char*>(blk[b - 1].data);
    if (b == 1) stride = d;
    uniform = uniform && d == stride && d > 0 && d % (long long)sizeof(T) == 0;
  }
  if (uniform) {
    const bool tr = blk[0].trans != 0;
    return gemm_dispatch<T>(h, precision, tr, false, blk[0].rows, k, blk[0].cols, alpha, static_cast<const T*>(blk[0].data),
                            blk[0].ld, stride / (long long)sizeof(T), V, ldv, blk[0].cols, beta, W, ldw, blk[0].rows,
                            nblocks, 0, nullptr, st);
  }
  long long r0 = 0, c0 = 0;
  for (int b = 0; b < nblocks; b++) {
    const long long m = blk[b].rows, n = blk[b].cols;
    const T* v = V + c0;
    T* w = W + r0;
    int rc = SCIMLOP_OK;
    switch (blk[b].kind & 0xFF) {
      case SCIMLOP_DENSE:
        rc = gemm_dispatch<T>(h, precision, blk[b].trans != 0, false, m, k, n, alpha, static_cast<const T*>(blk[b].data),
                              blk[b].ld, 0, v, ldv, 0, beta, w, ldw, 0, 1, 0, nullptr, st);
        break;
      case SCIMLOP_DIAG:
      case SCIMLOP_BDIAG:
        rc = diag_mul_t<T>(h, m, k, blk[b].data, (blk[b].kind & 0xFF) == SCIMLOP_DIAG ? 0 : blk[b].ld, v, ldv, w, ldw, &alpha,
                           &beta, st);
        break;
      case SCIMLOP_IDENTITY:
        rc = eltwise_axpby<T>(h, m, k, v, ldv, w, ldw, alpha, beta, st);
        break;
      case SCIMLOP_NULL:
        rc = eltwise_axpby<T>(h, m, k, nullptr, 0, w, ldw, T(0), beta, st);
        break;
      default:
        rc = scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "blockdiag: block %d has unsupported kind %d", b, blk[b].kind);
    }
    if (rc) return rc;
    r0 += m;
    c0 += n;
  }
  return SCIMLOP_OK;
}

}  // namespace

extern "C" int scimlop_blockdiag_mul(scimlop_handle_t h, int dtype, int precision, int nblocks,
                                     const scimlop_factor_t* blocks, const void* V, int64_t ldv, void* W, int64_t ldw,
                                     int64_t k, const void* alpha, const void* beta, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (int rc = check_dtype(h, dtype)) return rc;
  if (nblocks < 1 || !blocks) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "blockdiag: no blocks");
  if (k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "blockdiag: negative k");
  long long rows = 0, cols = 0;
  for (int b = 0; b < nblocks; b++) {
    const int kind = blocks[b].kind & 0xFF;
    if (blocks[b].rows < 0 || blocks[b].cols < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "blockdiag: negative block size");
    if (kind != SCIMLOP_DENSE && blocks[b].rows != blocks[b].cols)
      return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "blockdiag: block %d is elementwise but not square", b);
    if ((kind == SCIMLOP_DENSE || kind == SCIMLOP_DIAG || kind == SCIMLOP_BDIAG) && !blocks[b].data)
      return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "blockdiag: block %d has no data", b);
    if (kind == SCIMLOP_DENSE && blocks[b].ld < (blocks[b].trans ? blocks[b].cols : blocks[b].rows))
      return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "blockdiag: block %d leading dimension too small", b);
    rows += blocks[b].rows;
    cols += blocks[b].cols;
  }
  if (k == 0 || rows == 0) return SCIMLOP_OK;
  if (!V || !W) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "blockdiag: V/W null");
  if (ldv < cols || ldw < rows) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "blockdiag: leading dimension smaller than the operator");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SCIMLOP_F64)
    return blockdiag_t<double>(h, precision, nblocks, blocks, (const double*)V, ldv, (double*)W, ldw, k,
                               scalar_or<double>(alpha, 1.0), scalar_or<double>(beta, 0.0), st);
  return blockdiag_t<float>(h, precision, nblocks, blocks, (const float*)V, ldv, (float*)W, ldw, k,
                            scalar_or<float>(alpha, 1.f), scalar_or<float>(beta, 0.f), st);
}

// ================================================================================================
// TensorProductOperator
// ================================================================================================
namespace {

constexpr int MAX_KRON_FACTORS = 8;

struct KronPlan {
  int nops;                // non-identity factors
  long long rows, cols;    // prod of factor rows / cols
  long long max_inter;     // largest intermediate (elements per k column count included)
  long long csr_stage;     // elements of the staging copy an innermost SCIMLOP_CSR factor gathers from (0: none)
};

int kron_plan(scimlop_context* c, int nfactors, const int64_t* dims, const int32_t* kinds, int64_t k, KronPlan* pl) {
  if (nfactors < 1 || nfactors > MAX_KRON_FACTORS)
    return scimlop_fail(c, SCIMLOP_ERR_BAD_SHAPE, "kron: nfactors must be in 1..%d (got %d)", MAX_KRON_FACTORS, nfactors);
  if (!dims || !kinds) return scimlop_fail(c, SCIMLOP_ERR_NULL_POINTER, "kron: dims/kinds null");
  if (k < 0) return scimlop_fail(c, SCIMLOP_ERR_BAD_SHAPE, "kron: negative k");
  pl->nops = 0; pl->rows = 1; pl->cols = 1; pl->max_inter = 0;
  for (int f = 0; f < nfactors; f++) {
    const int64_t m = dims[2 * f], n = dims[2 * f + 1];
    if (m < 0 || n < 0) return scimlop_fail(c, SCIMLOP_ERR_BAD_SHAPE, "kron: negative factor size");
    if (kinds[f] == SCIMLOP_IDENTITY || kinds[f] == SCIMLOP_DIAG || kinds[f] == SCIMLOP_BANDED) {
      if (m != n) return scimlop_fail(c, SCIMLOP_ERR_BAD_SHAPE, "kron: identity/diagonal/banded factor %d must be square", f);
    } else if (kinds[f] == SCIMLOP_CSR) {
      // general sparse factor: any shape
    } else if ((kinds[f] & ~SCIMLOP_TRANS) != SCIMLOP_DENSE) {
      return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "kron: factor kind %d not supported inside a tensor product", kinds[f]);
    }
    if (kinds[f] != SCIMLOP_IDENTITY) pl->nops++;
    pl->rows *= m; pl->cols *= n;
  }
  // an innermost sparse factor gathers from a row-contiguous copy of the input (csrc/sparse.cu): cols * k elements
  pl->csr_stage = (kinds[nfactors - 1] == SCIMLOP_CSR) ? pl->cols * k : 0;
  // intermediates: after applying modes nfactors-1 .. p the tensor has prod_{q>=p} m_q * prod_{q<p} n_q * k entries
  long long done = 0;
  for (int p = nfactors - 1; p >= 0; p--) {
    if (kinds[p] == SCIMLOP_IDENTITY) continue;
    done++;
    if (done == pl->nops) break;  // the last op writes W
    long long e = k;
    for (int q = 0; q < nfactors; q++) e *= (q >= p) ? dims[2 * q] : dims[2 * q + 1];
    pl->max_inter = std::max(pl->max_inter, e);
  }
  return SCIMLOP_OK;
}

template <typename T>
int kron_mul_t(scimlop_context* c, int precision, int nfactors, const void* const* factors, const int64_t* dims,
               const int64_t* lds, const int32_t* kinds, const T* V, T* W, long long k, T alpha, T beta,
               void* workspace, size_t ws_bytes, const KronPlan& pl, cudaStream_t st, const PeerArg& g_peers) {
  if (pl.nops == 0) return eltwise_axpby<T>(c, pl.rows, k, V, pl.cols, W, pl.rows, alpha, beta, st);
  const size_t buf_bytes = align_up((size_t)pl.max_inter * sizeof(T), 256);
  const int nbuf = pl.nops >= 3 ? 2 : (pl.nops == 2 ? 1 : 0);
  if (nbuf > 0 && (!workspace || ws_bytes < buf_bytes * nbuf))
    return scimlop_fail(c, SCIMLOP_ERR_WORKSPACE, "kron: workspace %zu B < required %zu B (call cache_operator first)",
                        ws_bytes, buf_bytes * nbuf);
  T* bufs[2] = {static_cast<T*>(workspace),
                reinterpret_cast<T*>(static_cast<char*>(workspace) + buf_bytes)};
  const T* cur = V;
  int done = 0, flip = 0;
  long long L = 1;  // product of rows of the modes already applied (faster indices)
  for (int p = nfactors - 1; p >= 0; p--) {
    const long long m = dims[2 * p], n = dims[2 * p + 1];
    if (kinds[p] == SCIMLOP_IDENTITY) { L *= m; continue; }
    long long R = k;  // product of the slower, not yet applied modes and the batch
    for (int q = 0; q < p; q++) R *= dims[2 * q + 1];
    done++;
    const bool last = (done == pl.nops);
    T* out = last ? W : bufs[flip];
    const T a = last ? alpha : T(1), bt = last ? beta : T(0);
    const bool tr = (kinds[p] & SCIMLOP_TRANS) != 0;
    int rc;
    c->grid_limit = last ? c->stage_limit[1] : (done == 1 ? c->stage_limit[0] : 0);
    if constexpr (std::is_same<T, float>::value) {
      // the two fastest modes, both small dense Float32 factors: ONE pass with the intermediate kept on chip
      if (L == 1 && p >= 1 && (kinds[p] & ~SCIMLOP_TRANS) == SCIMLOP_DENSE && (kinds[p - 1] & ~SCIMLOP_TRANS) == SCIMLOP_DENSE &&
          !g_peers.active()) {
        const long long m2 = dims[2 * (p - 1)], n2 = dims[2 * (p - 1) + 1];
        const long long slabs = R / n2;
        const bool last2 = (done + 1 == pl.nops);
        float* out2 = last2 ? W : bufs[flip];
        if (kron2_f16_eligible(precision, m, n, m2, n2, slabs, cur, out2)) {
          rc = launch_kron2_f16(c, (int)m, (int)n, (int)m2, (int)n2, slabs, static_cast<const float*>(factors[p]), lds[p], tr,
                                static_cast<const float*>(factors[p - 1]), lds[p - 1], (kinds[p - 1] & SCIMLOP_TRANS) != 0, cur,
                                out2, last2 ? alpha : 1.f, last2 ? beta : 0.f, precision == SCIMLOP_PREC_BF16, st);
          if (rc) return rc;
          if (c->split_stream && c->stage_event && !last2) {
            // two-pass pipeline (kron_mul_impl_peers): the remaining pass of this column chunk runs on the side stream,
            // beside the fused pass of the next chunk on the caller's
            SCIMLOP_CUDA(c, cudaEventRecord(c->stage_event, st));
            SCIMLOP_CUDA(c, cudaStreamWaitEvent(c->split_stream, c->stage_event, 0));
            st = c->split_stream;
          }
          done++;
          cur = out2;
          flip ^= 1;
          L = m * m2;
          p--;
          continue;
        }
      }
    }
    if constexpr (std::is_same<T, double>::value) {
      // small problem, two small dense factors: one launch, intermediate in shared memory
      if (L == 1 && p >= 1 && (kinds[p] & ~SCIMLOP_TRANS) == SCIMLOP_DENSE && (kinds[p - 1] & ~SCIMLOP_TRANS) == SCIMLOP_DENSE &&
          !g_peers.active()) {
        const long long m2 = dims[2 * (p - 1)], n2 = dims[2 * (p - 1) + 1];
        const long long slabs = n2 > 0 ? R / n2 : 0;
        if (kron2_f64_small_eligible(c, m, n, m2, n2, slabs)) {
          const bool last2 = (done + 1 == pl.nops);
          double* out2 = last2 ? W : bufs[flip];
          rc = launch_kron2_f64_small(c, (int)m, (int)n, (int)m2, (int)n2, slabs, static_cast<const double*>(factors[p]), lds[p], tr,
                                      static_cast<const double*>(factors[p - 1]), lds[p - 1], (kinds[p - 1] & SCIMLOP_TRANS) != 0,
                                      cur, out2, last2 ? alpha : 1.0, last2 ? beta : 0.0, st);
          if (rc) return rc;
          done++;
          cur = out2;
          flip ^= 1;
          L = m * m2;
          p--;
          continue;
        }
      }
    }
    if (kinds[p] == SCIMLOP_CSR) {
      // mode product with a general sparse matrix: gather kernels of csrc/sparse.cu                [L2 / HBM-bound]
      if (g_peers.active() && last) return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "kron: peer targets need a dense outermost factor");
      char* stage = nullptr;
      size_t stage_bytes = 0;
      if (L == 1 && workspace && ws_bytes >= buf_bytes * nbuf + (size_t)pl.csr_stage * sizeof(T)) {
        stage = static_cast<char*>(workspace) + buf_bytes * nbuf;
        stage_bytes = ws_bytes - buf_bytes * nbuf;
      }
      rc = scimlop_csr_mode_impl(c, std::is_same<T, double>::value ? SCIMLOP_F64 : SCIMLOP_F32, factors[p], L, m, n, R, cur, out, &a, &bt,
                                 stage, stage_bytes, st);
    } else if (kinds[p] == SCIMLOP_BANDED) {
      // mode product with a banded matrix: a 1-D stencil along index (i / L) % n      [HBM-bound]
      const long long total = L * n * R;
      const int b = (int)lds[p];
      if (b < 0 || b > band::MAXB) return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "kron: half-bandwidth %d not in 0..%d", b, band::MAXB);
      const long long blocks = std::min<long long>((total + band::THREADS - 1) / band::THREADS, (long long)c->sm_count * 32);
      band::band_mode_kernel<T><<<(unsigned)std::max<long long>(1, blocks), band::THREADS, 0, st>>>(
          total, L, n, b, static_cast<const T*>(factors[p]), cur, out, a, bt);
      c->launches++;
      rc = (cudaGetLastError() == cudaSuccess) ? SCIMLOP_OK : scimlop_fail(c, SCIMLOP_ERR_CUDA, "band_mode_kernel launch failed");
    } else if (kinds[p] == SCIMLOP_DIAG) {
      // mode product with a diagonal: scale along index (i / L) % n of every column of the L*n x R view
      elt::Params<T> ep;
      memset(&ep, 0, sizeof(ep));
      ep.n = L * n; ep.k = R; ep.V = cur; ep.ldv = L * n; ep.W = out; ep.ldw = L * n; ep.alpha = a; ep.beta = bt;
      ep.nterms = 1; ep.coef[0] = T(1); ep.nvec[0] = 1; ep.nfac[0] = 1;
      ep.fac[0][0].ptr = static_cast<const T*>(factors[p]); ep.fac[0][0].ld = 0; ep.fac[0][0].div = L; ep.fac[0][0].mod = n;
      rc = launch_eltwise(c, ep, st);
    } else if (L == 1) {
      // innermost applied mode: out(m x R) = op(F)(m x n) * cur(n x R)    [src/tensor.jl:568,603]
      rc = gemm_dispatch<T>(c, precision, tr, false, m, R, n, a, static_cast<const T*>(factors[p]), lds[p], 0, cur,
                            n, 0, bt, out, m, 0, 1, 0, nullptr, st, last ? g_peers : NO_PEERS);
    } else {
      // out_r(L x m) = cur_r(L x n) * op(F)^T for every slab r           [src/tensor.jl:778-788, no permutedims]
      rc = gemm_dispatch<T>(c, precision, false, !tr, L, m, n, a, cur, L, L * n, static_cast<const T*>(factors[p]),
                            lds[p], 0, bt, out, L, L * m, R, 0, nullptr, st, last ? g_peers : NO_PEERS);
    }
    c->grid_limit = 0;
    if (rc) return rc;
    if (done == 1 && c->stage_event) SCIMLOP_CUDA(c, cudaEventRecord(c->stage_event, st));
    cur = out;
    flip ^= 1;
    L *= m;
  }
  return SCIMLOP_OK;
}


// ---- complex element types (ComplexF32 / ComplexF64: test/basic.jl:405-426, test/Downstream/alloccheck.jl:37) ---------------
// Julia's Complex arrays interleave (re, im).  A complex mode product is four real GEMMs on PLANAR data
//     Yr = op(Fr) Xr - s op(Fi) Xi,   Yi = s op(Fi) Xr + op(Fr) Xi          (s = -1 for the lazy adjoint F^H)
// -- exactly the 8 m n k real flops of a complex product -- so the planner de-interleaves V once, runs every mode of the
// Kronecker product as 4 tensor-core GEMMs between two planar ping-pong buffers (the same strided-batched shapes as the
// real path: no permutedims), and re-interleaves into W with the complex alpha / beta applied in that last pass.
template <typename T>
__global__ void cplx_split_kernel(long long rows, long long cols, const T* __restrict__ Z, long long ldz, T* __restrict__ re,
                                  T* __restrict__ im) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i % rows, c = i / rows;
    const T* z = Z + 2 * (r + c * ldz);
    re[i] = z[0];
    im[i] = z[1];
  }
}
template <typename T>
__global__ void cplx_merge_kernel(long long rows, long long cols, const T* __restrict__ re, const T* __restrict__ im, T* W,
                                  long long ldw, T ar, T ai, T br, T bi, int readw) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i % rows, c = i / rows;
    T* w = W + 2 * (r + c * ldw);
    const T yr = re[i], yi = im[i];
    T outr = ar * yr - ai * yi, outi = ar * yi + ai * yr;
    if (readw) {
      const T wr = w[0], wi = w[1];
      outr += br * wr - bi * wi;
      outi += br * wi + bi * wr;
    }
    w[0] = outr;
    w[1] = outi;
  }
}

struct CplxLayout {
  size_t plane;            // bytes of one plane of one ping-pong buffer
  size_t fac_off[MAX_KRON_FACTORS];
  size_t total;
};
inline CplxLayout cplx_layout(int nfactors, const int64_t* dims, const int32_t* kinds, long long k, const KronPlan& pl, size_t es) {
  CplxLayout L;
  long long maxe = std::max(pl.rows, pl.cols) * k;
  maxe = std::max(maxe, pl.max_inter);
  L.plane = align_up((size_t)maxe * es, 256);
  size_t off = 4 * L.plane;                                 // two buffers x (re, im)
  for (int f = 0; f < nfactors; f++) {
    L.fac_off[f] = off;
    if (kinds[f] != SCIMLOP_IDENTITY) off += 2 * align_up((size_t)dims[2 * f] * (size_t)dims[2 * f + 1] * es, 256);
  }
  L.total = off;
  return L;
}

template <typename T>
int kron_mul_complex_t(scimlop_context* c, int precision, int nfactors, const void* const* factors, const int64_t* dims,
                       const int64_t* lds, const int32_t* kinds, const T* V, long long ldv, T* W, long long ldw, long long k,
                       const T* alpha, const T* beta, void* workspace, size_t ws_bytes, const KronPlan& pl, cudaStream_t st) {
  for (int f = 0; f < nfactors; f++)
    if (kinds[f] != SCIMLOP_IDENTITY && (kinds[f] & ~SCIMLOP_TRANS) != SCIMLOP_DENSE)
      return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "kron (complex): factor %d: only dense and identity factors", f);
  const CplxLayout lay = cplx_layout(nfactors, dims, kinds, k, pl, sizeof(T));
  if (!workspace || ws_bytes < lay.total)
    return scimlop_fail(c, SCIMLOP_ERR_WORKSPACE, "kron (complex): workspace %zu B < required %zu B", ws_bytes, lay.total);
  char* base = static_cast<char*>(workspace);
  T* buf[2][2] = {{reinterpret_cast<T*>(base), reinterpret_cast<T*>(base + lay.plane)},
                  {reinterpret_cast<T*>(base + 2 * lay.plane), reinterpret_cast<T*>(base + 3 * lay.plane)}};
  auto grid_of = [&](long long total) {
    return (unsigned)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)c->sm_count * 16));
  };
  const T ar = alpha ? alpha[0] : T(1), ai = alpha ? alpha[1] : T(0), br = beta ? beta[0] : T(0), bi = beta ? beta[1] : T(0);
  const int readw = (br != T(0) || bi != T(0)) ? 1 : 0;
  // V -> planar
  cplx_split_kernel<T><<<grid_of(pl.cols * k), 256, 0, st>>>(pl.cols, k, V, ldv, buf[0][0], buf[0][1]);
  c->launches++;
  int cur = 0;
  long long L = 1;
  for (int p = nfactors - 1; p >= 0; p--) {
    const long long m = dims[2 * p], n = dims[2 * p + 1];
    if (kinds[p] == SCIMLOP_IDENTITY) { L *= m; continue; }
    long long R = k;
    for (int q = 0; q < p; q++) R *= dims[2 * q + 1];
    const bool tr = (kinds[p] & SCIMLOP_TRANS) != 0;
    const long long srows = tr ? n : m, scols = tr ? m : n;          // stored shape
    T* Fr = reinterpret_cast<T*>(base + lay.fac_off[p]);
    T* Fi = reinterpret_cast<T*>(base + lay.fac_off[p] + align_up((size_t)m * (size_t)n * sizeof(T), 256));
    cplx_split_kernel<T><<<grid_of(srows * scols), 256, 0, st>>>(srows, scols, static_cast<const T*>(factors[p]), lds[p], Fr, Fi);
    c->launches++;
    const T sgn = tr ? T(-1) : T(1);                                  // lazy adjoint: conjugate transpose
    const T *Xr = buf[cur][0], *Xi = buf[cur][1];
    T *Yr = buf[cur ^ 1][0], *Yi = buf[cur ^ 1][1];
    auto gemm = [&](const T* F, const T* X, T* Y, T a, T b) -> int {
      if (L == 1)   // innermost applied mode: Y(m x R) = op(F)(m x n) X(n x R)
        return gemm_dispatch<T>(c, precision, tr, false, m, R, n, a, F, srows, 0, X, n, 0, b, Y, m, 0, 1, 0, nullptr, st, NO_PEERS);
      return gemm_dispatch<T>(c, precision, false, !tr, L, m, n, a, X, L, L * n, F, srows, 0, b, Y, L, L * m, R, 0, nullptr, st, NO_PEERS);
    };
    int rc = gemm(Fr, Xr, Yr, T(1), T(0));
    if (!rc) rc = gemm(Fi, Xi, Yr, -sgn, T(1));
    if (!rc) rc = gemm(Fi, Xr, Yi, sgn, T(0));
    if (!rc) rc = gemm(Fr, Xi, Yi, T(1), T(1));
    if (rc) return rc;
    cur ^= 1;
    L *= m;
  }
  cplx_merge_kernel<T><<<grid_of(pl.rows * k), 256, 0, st>>>(pl.rows, k, buf[cur][0], buf[cur][1], W, ldw, ar, ai, br, bi, readw);
  c->launches++;
  SCIMLOP_CUDA(c, cudaGetLastError());
  return SCIMLOP_OK;
}

}  // namespace

extern "C" int scimlop_kron_workspace_bytes(int dtype, int nfactors, const int64_t* dims, const int32_t* kinds,
                                            int64_t k, size_t* bytes) {
  if (!bytes) return SCIMLOP_ERR_NULL_POINTER;
  if (dtype == SCIMLOP_C64 || dtype == SCIMLOP_C128) {
    KronPlan plc;
    if (int rc = kron_plan(nullptr, nfactors, dims, kinds, k, &plc)) return rc;
    *bytes = cplx_layout(nfactors, dims, kinds, k, plc, dtype == SCIMLOP_C128 ? 8 : 4).total;
    return SCIMLOP_OK;
  }
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32) return SCIMLOP_ERR_UNSUPPORTED;
  KronPlan pl;
  if (int rc = kron_plan(nullptr, nfactors, dims, kinds, k, &pl)) return rc;
  const int nbuf = pl.nops >= 3 ? 2 : (pl.nops == 2 ? 1 : 0);
  *bytes = nbuf * align_up((size_t)pl.max_inter * elem_size(dtype), 256) + align_up((size_t)pl.csr_stage * elem_size(dtype), 256);
  return SCIMLOP_OK;
}

static int kron_mul_impl_peers(scimlop_context* h, int dtype, int precision, int nfactors, const void* const* factors,
                               const int64_t* dims, const int64_t* lds, const int32_t* kinds, const void* V, int64_t ldv,
                               void* W, int64_t ldw, int64_t k, const void* alpha, const void* beta, void* workspace,
                               size_t workspace_bytes, cudaStream_t st, const PeerArg& peers) {
  const bool is_complex = dtype == SCIMLOP_C64 || dtype == SCIMLOP_C128;
  if (!is_complex)
    if (int rc = check_dtype(h, dtype)) return rc;
  if (is_complex && peers.active()) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "peer targets: real element types only");
  KronPlan pl;
  if (int rc = kron_plan(h, nfactors, dims, kinds, k, &pl)) return rc;
  if (!factors || !lds) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron: factors/lds null");
  for (int f = 0; f < nfactors; f++) {
    if (kinds[f] == SCIMLOP_IDENTITY) continue;
    if (!factors[f]) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron: factor %d is null", f);
    if ((kinds[f] & ~SCIMLOP_TRANS) == SCIMLOP_DENSE &&
        lds[f] < std::max<int64_t>(1, (kinds[f] & SCIMLOP_TRANS) ? dims[2 * f + 1] : dims[2 * f]))
      return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kron: ld of factor %d smaller than its stored rows", f);
  }
  if (k == 0 || pl.rows == 0) return SCIMLOP_OK;
  if (!W || (!V && pl.cols > 0)) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron: V/W null");
  // the reshape views of src/tensor.jl:558,593 need contiguous columns
  if (ldv != pl.cols || ldw != pl.rows)
    return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE,
                        "kron: V and W must be contiguous (ldv=%lld, cols=%lld; ldw=%lld, rows=%lld)", (long long)ldv,
                        pl.cols, (long long)ldw, pl.rows);
  if (is_complex) {
    if (pl.cols == 0) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "kron (complex): empty contraction");
    if (dtype == SCIMLOP_C128)
      return kron_mul_complex_t<double>(h, SCIMLOP_PREC_DEFAULT, nfactors, factors, dims, lds, kinds, (const double*)V, ldv,
                                        (double*)W, ldw, k, (const double*)alpha, (const double*)beta, workspace,
                                        workspace_bytes, pl, st);
    return kron_mul_complex_t<float>(h, precision, nfactors, factors, dims, lds, kinds, (const float*)V, ldv, (float*)W, ldw, k,
                                     (const float*)alpha, (const float*)beta, workspace, workspace_bytes, pl, st);
  }
  if (pl.cols == 0) {  // empty contraction: W = beta * W
    if (dtype == SCIMLOP_F64)
      return eltwise_axpby<double>(h, pl.rows, k, nullptr, 0, (double*)W, ldw, 0.0, scalar_or<double>(beta, 0.0), st);
    return eltwise_axpby<float>(h, pl.rows, k, nullptr, 0, (float*)W, ldw, 0.f, scalar_or<float>(beta, 0.f), st);
  }
  // the fused compute + all-gather entries keep ONE summation order in every stage (their last stage cannot take the
  // split-K tail launch, so no stage does: results are bit-identical to the SCIMLOP_OPT_TAIL_SPLIT = 0 schedule)
  const int saved_tail_split = h->tail_split;
  if (peers.active()) h->tail_split = 0;
  // Column-chunk pipeline of the fused all-gather: the last stage of a chunk is bound by NVLink ingress, not by the SMs,
  // so it runs on a share of them while the first stage of the NEXT chunk computes on the rest (second stream), instead
  // of the whole first stage running before any store is on the wire.  Columns are independent: same numbers, bit for bit.
  // The whole problem must have >= sm_count tiles in its first stage (it then runs the plain persistent kernel, and so
  // do the chunks: no_small_splitk), a chunk at least 3/4 of that; chunk offsets keep the 256 B alignment of the
  // workspace and 16 B of V / W.
  int nch = 0;
  double ingress_us = 0.0, stage2_us = 1.0;
  long long t1c = 0, t2c = 0;   // tiles of one chunk in the first / the last stage
  if (peers.active() && h->peer_pipeline && dtype == SCIMLOP_F64 && nfactors == 2 && (kinds[0] & ~SCIMLOP_TRANS) == SCIMLOP_DENSE &&
      (kinds[1] & ~SCIMLOP_TRANS) == SCIMLOP_DENSE && scalar_or<double>(beta, 0.0) == 0.0 && (pl.max_inter / k) % 32 == 0 &&
      pl.rows % 2 == 0 && pl.cols % 2 == 0) {
    const long long mt = (dims[2] + 127) / 128;
    if (mt * ((k * dims[1] + 127) / 128) >= h->sm_count)
      for (int cand = h->peer_chunks ? h->peer_chunks : 2; cand >= 2 && !nch; cand -= 2)
        if (k % cand == 0 && (h->peer_chunks || mt * ((k / cand * dims[1] + 127) / 128) >= h->sm_count)) nch = cand;
    if (nch) {
      t1c = mt * ((k / nch * dims[1] + 127) / 128);
      t2c = mt * ((dims[0] + 127) / 128) * (k / nch);
    }
    // worth it only where the gather, not the arithmetic, bounds the last stage: bytes this GPU receives at ~600 GB/s
    // (NVLink 5 ingress as NCCL's all-gather reaches it here) against the stage's flops at ~35 TFLOP/s (DMMA, measured)
    const int npeers = peers.n > 0 ? peers.n : std::max(0, h->nranks - 1);
    ingress_us = (double)npeers * (double)pl.rows * (double)k * 8.0 / 6.0e5;
    stage2_us = 2.0 * (double)dims[1] * (double)pl.rows * (double)k / 35.0e6;
    // measured (profiles/r02_peer_pipeline_sweep*.jsonl): ratio 3.2 (8 ranks) -13 %, 1.37 (4 ranks) -9 %, 0.45 (2 ranks) +21 %
    if (h->peer_pipeline == 1 && ingress_us < 1.25 * stage2_us) nch = 0;
  }
  if (nch) {
    if (int rc0 = scimlop_ensure_side_stream(h)) { h->tail_split = saved_tail_split; return rc0; }
    const int nsm = std::max(1, h->sm_count - h->sm_reserve);
    // SMs of the last stage of a non-final chunk.  The next chunk's first stage is on the critical path (its last stage
    // starts when it ends and then shares the wire), so: fewest waves for that first stage on the remaining SMs, among
    // the splits whose own waves still fit the gather time of a chunk; of those the smallest share + 4 SMs of slack.
    // Measured on 8 GPUs (profiles/r02_peer_pipeline_sweep.jsonl): 2 chunks x 56 SMs 0.996 ms, 64..100 SMs 1.06-1.07,
    // 4 chunks 1.09-1.14, unchunked 1.14.
    int g2 = nsm / 3;
    {
      const double tile_us = stage2_us / std::max(1.0, (double)(t2c * nch) / nsm);
      const double wire = ingress_us / nch / std::max(tile_us, 1e-3);
      const int lo = std::max(1, nsm / 4), hi = std::max(lo, nsm / 2);
      auto w2 = [&](int s) { return (double)((t2c + s - 1) / s); };
      auto w1 = [&](int s) { return (double)((t1c + (nsm - s) - 1) / (nsm - s)); };
      const double w2_ok = std::max(wire, w2(hi));
      double best = 1e30;
      for (int s = lo; s <= hi; s++)
        if (w2(s) <= w2_ok) best = std::min(best, w1(s));
      int first = 0, last = 0;
      for (int s = lo; s <= hi; s++)
        if (w2(s) <= w2_ok && w1(s) == best) { if (!first) first = s; last = s; }
      if (first) g2 = std::min(first + 4, last);
    }
    if (h->peer_stage2_sms > 0) g2 = std::min(h->peer_stage2_sms, nsm - 1);
    const int g1 = nsm - g2;
    h->no_small_splitk = 1;
    const long long inter_per_col = pl.max_inter / k;
    const double a = scalar_or<double>(alpha, 1.0);
    SCIMLOP_CUDA(h, cudaEventRecord(h->ev_pipe[8], st));
    SCIMLOP_CUDA(h, cudaStreamWaitEvent(h->s_side, h->ev_pipe[8], 0));
    int rc = SCIMLOP_OK;
    long long c0 = 0;
    for (int i = 0; i < nch && rc == SCIMLOP_OK; i++) {
      const long long nc = (k * (i + 1)) / nch - c0;
      cudaStream_t si = (i % 2 == 0) ? st : h->s_side;
      if (i > 0) SCIMLOP_CUDA(h, cudaStreamWaitEvent(si, h->ev_pipe[i - 1], 0));   // S1(i) starts when S1(i-1) is done
      KronPlan plc;
      if ((rc = kron_plan(h, nfactors, dims, kinds, nc, &plc))) break;
      PeerArg pc;
      pc.n = peers.n;
      for (int q = 0; q < peers.n; q++) pc.ptr[q] = static_cast<double*>(peers.ptr[q]) + c0 * pl.rows;
      if (peers.mc) pc.mc = static_cast<double*>(peers.mc) + c0 * pl.rows;
      h->stage_event = h->ev_pipe[i];
      h->stage_limit[0] = i == 0 ? 0 : g1;
      h->stage_limit[1] = i == nch - 1 ? 0 : g2;
      rc = kron_mul_t<double>(h, precision, nfactors, factors, dims, lds, kinds, static_cast<const double*>(V) + c0 * pl.cols,
                              static_cast<double*>(W) + c0 * pl.rows, nc, a, 0.0 /* checked above */,
                              static_cast<char*>(workspace) + (size_t)c0 * (size_t)inter_per_col * 8,
                              workspace_bytes - (size_t)c0 * (size_t)inter_per_col * 8, plc, si, pc);
      c0 += nc;
    }
    h->stage_event = nullptr;
    h->no_small_splitk = 0;
    h->stage_limit[0] = h->stage_limit[1] = 0;
    h->grid_limit = 0;
    h->tail_split = saved_tail_split;
    if (rc) return rc;
    SCIMLOP_CUDA(h, cudaEventRecord(h->ev_pipe[9], h->s_side));
    SCIMLOP_CUDA(h, cudaStreamWaitEvent(st, h->ev_pipe[9], 0));
    return SCIMLOP_OK;
  }
  // Two-pass pipeline for three small dense Float32 factors (config 3): the fused two-mode pass is bound by the shared-
  // memory pipe, the third mode's pass by HBM.  Over column chunks, pass 2 of chunk i runs on the side stream on a share
  // of the SMs beside pass 1 of chunk i + 1 on the rest.  Columns are independent: same numbers as the unchunked call.
  if (!peers.active() && h->mode_pipeline >= 2 && dtype == SCIMLOP_F32 && nfactors == 3 && k % h->mode_pipeline == 0 &&
      (kinds[0] & ~SCIMLOP_TRANS) == SCIMLOP_DENSE && (kinds[1] & ~SCIMLOP_TRANS) == SCIMLOP_DENSE &&
      (kinds[2] & ~SCIMLOP_TRANS) == SCIMLOP_DENSE) {
    const int nch = h->mode_pipeline;
    const long long kc = k / nch;
    KronPlan plc;
    if (int rc0 = kron_plan(h, nfactors, dims, kinds, kc, &plc)) { h->tail_split = saved_tail_split; return rc0; }
    const size_t bufc = align_up((size_t)plc.max_inter * 4, 256);
    const long long slabs_c = kc * dims[1];      // slabs of one chunk's fused pass (inner two modes)
    if (kron2_f16_eligible(precision, dims[4], dims[5], dims[2], dims[3], slabs_c, (const float*)V, (const float*)workspace) &&
        workspace && workspace_bytes >= 2 * bufc * (size_t)nch && (pl.rows * kc) % 4 == 0 && (pl.cols * kc) % 4 == 0) {
      if (int rc0 = scimlop_ensure_side_stream(h)) { h->tail_split = saved_tail_split; return rc0; }
      const int nsm = std::max(1, h->sm_count - h->sm_reserve);
      const int g2 = h->peer_stage2_sms > 0 ? std::min(h->peer_stage2_sms, nsm - 1) : (nsm * 2) / 5, g1 = nsm - g2;
      const float a = scalar_or<float>(alpha, 1.f), b = scalar_or<float>(beta, 0.f);
      SCIMLOP_CUDA(h, cudaEventRecord(h->ev_pipe[8], st));
      SCIMLOP_CUDA(h, cudaStreamWaitEvent(h->s_side, h->ev_pipe[8], 0));
      int rc = SCIMLOP_OK;
      for (int i = 0; i < nch && rc == SCIMLOP_OK; i++) {
        h->split_stream = h->s_side;
        h->stage_event = h->ev_pipe[i];
        h->stage_limit[0] = i == 0 ? 0 : g1;
        h->stage_limit[1] = i == nch - 1 ? 0 : g2;
        rc = kron_mul_t<float>(h, precision, nfactors, factors, dims, lds, kinds, static_cast<const float*>(V) + (size_t)i * kc * pl.cols,
                               static_cast<float*>(W) + (size_t)i * kc * pl.rows, kc, a, b,
                               static_cast<char*>(workspace) + (size_t)i * 2 * bufc, 2 * bufc, plc, st, NO_PEERS);
      }
      h->split_stream = nullptr;
      h->stage_event = nullptr;
      h->stage_limit[0] = h->stage_limit[1] = 0;
      h->grid_limit = 0;
      h->tail_split = saved_tail_split;
      if (rc) return rc;
      SCIMLOP_CUDA(h, cudaEventRecord(h->ev_pipe[9], h->s_side));
      SCIMLOP_CUDA(h, cudaStreamWaitEvent(st, h->ev_pipe[9], 0));
      return SCIMLOP_OK;
    }
  }
  int rc;
  if (dtype == SCIMLOP_F64)
    rc = kron_mul_t<double>(h, precision, nfactors, factors, dims, lds, kinds, (const double*)V, (double*)W, k,
                            scalar_or<double>(alpha, 1.0), scalar_or<double>(beta, 0.0), workspace, workspace_bytes,
                            pl, st, peers);
  else
    rc = kron_mul_t<float>(h, precision, nfactors, factors, dims, lds, kinds, (const float*)V, (float*)W, k,
                           scalar_or<float>(alpha, 1.f), scalar_or<float>(beta, 0.f), workspace, workspace_bytes, pl,
                           st, peers);
  h->tail_split = saved_tail_split;
  return rc;
}

int scimlop_kron_mul_impl(scimlop_context* h, int dtype, int precision, int nfactors, const void* const* factors,
                          const int64_t* dims, const int64_t* lds, const int32_t* kinds, const void* V, int64_t ldv,
                          void* W, int64_t ldw, int64_t k, const void* alpha, const void* beta, void* workspace,
                          size_t workspace_bytes, cudaStream_t st) {
  return kron_mul_impl_peers(h, dtype, precision, nfactors, factors, dims, lds, kinds, V, ldv, W, ldw, k, alpha, beta,
                             workspace, workspace_bytes, st, NO_PEERS);
}

extern "C" int scimlop_kron_mul(scimlop_handle_t h, int dtype, int precision, int nfactors,
                                const void* const* factors, const int64_t* dims, const int64_t* lds,
                                const int32_t* kinds, const void* V, int64_t ldv, void* W, int64_t ldw, int64_t k,
                                const void* alpha, const void* beta, void* workspace, size_t workspace_bytes,
                                void* stream) {
  if (int rc = check_handle(h)) return rc;
  scimlop_device_guard guard(h->device);
  return scimlop_kron_mul_impl(h, dtype, precision, nfactors, factors, dims, lds, kinds, V, ldv, W, ldw, k, alpha,
                               beta, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int scimlop_kron_mul_peers(scimlop_handle_t h, int dtype, int precision, int nfactors,
                                      const void* const* factors, const int64_t* dims, const int64_t* lds,
                                      const int32_t* kinds, const void* V, int64_t k, void* W_local,
                                      void* const* peer_targets, int npeers, const void* alpha, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (dtype != SCIMLOP_F64) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "kron_mul_peers: Float64 only");
  if (npeers < 0 || npeers > 7) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kron_mul_peers: 0..7 peers");
  if (npeers > 0 && !peer_targets) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_mul_peers: peer_targets null");
  if (!dims || !kinds || nfactors < 1 || nfactors > 8) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_mul_peers: bad factors");
  if (kinds[0] != SCIMLOP_DENSE && kinds[0] != (SCIMLOP_DENSE | SCIMLOP_TRANS))
    // the stage that writes W is the outermost factor's mode product; it must be a GEMM to carry the peer stores
    return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "kron_mul_peers: the outermost factor must be dense");
  long long rows = 1, cols = 1;
  for (int f = 0; f < nfactors; f++) { rows *= dims[2 * f]; cols *= dims[2 * f + 1]; }
  scimlop_device_guard guard(h->device);
  PeerArg peers;
  peers.n = npeers;
  for (int q = 0; q < npeers; q++) peers.ptr[q] = peer_targets[q];
  return kron_mul_impl_peers(h, dtype, precision, nfactors, factors, dims, lds, kinds, V, cols, W_local, rows, k, alpha,
                             nullptr, workspace, workspace_bytes, static_cast<cudaStream_t>(stream), peers);
}

extern "C" int scimlop_kron_mul_mc(scimlop_handle_t h, int dtype, int precision, int nfactors,
                                   const void* const* factors, const int64_t* dims, const int64_t* lds,
                                   const int32_t* kinds, const void* V, int64_t k, void* W_local, void* W_multicast,
                                   const void* alpha, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (dtype != SCIMLOP_F64) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "kron_mul_mc: Float64 only");
  if (!W_multicast) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_mul_mc: multicast pointer null");
  if (!dims || !kinds || nfactors < 1 || nfactors > 8) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_mul_mc: bad factors");
  if (kinds[0] != SCIMLOP_DENSE && kinds[0] != (SCIMLOP_DENSE | SCIMLOP_TRANS))
    return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "kron_mul_mc: the outermost factor must be dense");
  long long rows = 1, cols = 1;
  for (int f = 0; f < nfactors; f++) { rows *= dims[2 * f]; cols *= dims[2 * f + 1]; }
  scimlop_device_guard guard(h->device);
  PeerArg peers;
  peers.mc = W_multicast;
  return kron_mul_impl_peers(h, dtype, precision, nfactors, factors, dims, lds, kinds, V, cols, W_local, rows, k, alpha,
                             nullptr, workspace, workspace_bytes, static_cast<cudaStream_t>(stream), peers);
}

extern "C" int scimlop_outer_mul_fast(scimlop_handle_t h, int dtype, int precision, void* W, const void* A,
                                      int64_t lda, const void* C, int64_t mi, int64_t mo, int64_t no, int64_t k,
                                      const void* alpha, const void* beta, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (int rc = check_dtype(h, dtype)) return rc;
  if (mi < 0 || mo < 0 || no < 0 || k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "outer_mul: negative size");
  if (mi == 0 || mo == 0 || k == 0) return SCIMLOP_OK;
  if (!W || (no > 0 && (!A || !C))) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "outer_mul: null operand");
  if (no > 0 && lda < mo) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "outer_mul: lda < mo");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // W_j(mi x mo) = alpha * C_j(mi x no) * A^T + beta * W_j, j = 1..k
  if (dtype == SCIMLOP_F64)
    return gemm_dispatch<double>(h, precision, false, true, mi, mo, no, scalar_or<double>(alpha, 1.0),
                                 (const double*)C, mi, mi * no, (const double*)A, lda, 0, scalar_or<double>(beta, 0.0),
                                 (double*)W, mi, mi * mo, k, 0, nullptr, st);
  return gemm_dispatch<float>(h, precision, false, true, mi, mo, no, scalar_or<float>(alpha, 1.f), (const float*)C, mi,
                              mi * no, (const float*)A, lda, 0, scalar_or<float>(beta, 0.f), (float*)W, mi, mi * mo, k,
                              0, nullptr, st);
}

template <typename T>
static int kronsum_t(scimlop_context* h, int precision, const T* A, int64_t lda, int64_t mo, const T* B, int64_t ldb,
                     int64_t mi, const T* V, T* W, int64_t k, T alpha, T beta, cudaStream_t st) {
  // I (x) B : W(mi x mo*k) = alpha * B * V + beta * W                         [src/tensor.jl:389,401 / K6]
  int rc = gemm_dispatch<T>(h, precision, false, false, mi, mo * k, mi, alpha, B, ldb, 0, V, mi, 0, beta, W, mi, 0, 1, 0,
                            nullptr, st);
  if (rc) return rc;
  // A (x) I : W_j += alpha * V_j * A^T                                        [src/tensor.jl:388,400 / K7+K2]
  return gemm_dispatch<T>(h, precision, false, true, mi, mo, mo, alpha, V, mi, mi * mo, A, lda, 0, T(1), W, mi, mi * mo,
                          k, 0, nullptr, st);
}

extern "C" int scimlop_kronsum_mul(scimlop_handle_t h, int dtype, int precision, const void* A, int64_t lda,
                                   int64_t mo, const void* B, int64_t ldb, int64_t mi, const void* V, int64_t ldv,
                                   void* W, int64_t ldw, int64_t k, const void* alpha, const void* beta,
                                   void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (int rc = check_dtype(h, dtype)) return rc;
  if (mo < 0 || mi < 0 || k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kronsum: negative size");
  if (mo == 0 || mi == 0 || k == 0) return SCIMLOP_OK;
  if (!A || !B || !V || !W) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kronsum: null operand");
  if (lda < mo || ldb < mi) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kronsum: leading dimension smaller than rows");
  if (ldv != mi * mo || ldw != mi * mo)
    return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kronsum: V and W must be contiguous (ld == mi*mo)");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SCIMLOP_F64)
    return kronsum_t<double>(h, precision, (const double*)A, lda, mo, (const double*)B, ldb, mi, (const double*)V,
                             (double*)W, k, scalar_or<double>(alpha, 1.0), scalar_or<double>(beta, 0.0), st);
  return kronsum_t<float>(h, precision, (const float*)A, lda, mo, (const float*)B, ldb, mi, (const float*)V, (float*)W,
                          k, scalar_or<float>(alpha, 1.f), scalar_or<float>(beta, 0.f), st);
}

template <typename T>
static int kronsum_banded_t(scimlop_context* h, const T* Bx, int bx, int64_t mo, const T* By, int by, int64_t mi, const T* V,
                            T* W, int64_t k, T alpha, T beta, cudaStream_t st) {
  if (bx == 1 && by == 1 && aligned16(V) && (mi * sizeof(T)) % 16 == 0 && mi >= 64 && mo >= 16 &&
      mi < 2147483647LL && mo < 2147483647LL && k < 2147483647LL) {
    // TMA-tiled 5-point stencil: halo tiles stream through a 3-deep bulk-copy ring
    CUtensorMap tmV;
    cuuint64_t gdim[3] = {(cuuint64_t)mi, (cuuint64_t)mo, (cuuint64_t)k};
    cuuint64_t gstr[2] = {(cuuint64_t)mi * sizeof(T), (cuuint64_t)mi * (cuuint64_t)mo * sizeof(T)};
    cuuint32_t box[3] = {band::TTI, band::TTO, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = h->encode_tiled(&tmV, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                                 const_cast<T*>(V), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUtensorMap tmW = tmV;
    const bool readw = beta != T(0);
    if (r == CUDA_SUCCESS && readw) {
      // the old W tile travels by TMA too (needs the same alignment of W as of V; else the other stencil paths)
      if (!aligned16(W)) r = CUDA_ERROR_INVALID_VALUE;
      else {
        cuuint32_t wbox[3] = {(cuuint32_t)band::Tri<T>::IN, band::TO_IN, 1};
        r = h->encode_tiled(&tmW, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, W, gdim,
                            gstr, wbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
    }
    if (r == CUDA_SUCCESS) {
      band::TriParams p;
      p.mi = mi; p.mo = mo; p.k = k;
      p.tiles_i = (int)((mi + band::Tri<T>::IN - 1) / band::Tri<T>::IN);
      p.tiles_o = (int)((mo + band::TO_IN - 1) / band::TO_IN);
      p.total_tiles = (long long)p.tiles_i * p.tiles_o * k;
      const int grid = (int)std::min<long long>(p.total_tiles, (long long)h->sm_count * (readw ? 1 : 2));
      if (!readw)
        band::band_kronsum_tri_tma_kernel<T, false><<<grid, band::TTHREADS, band::tri_smem_bytes<T>(), st>>>(tmV, tmW, p, Bx, By, W, alpha, beta);
      else
        band::band_kronsum_tri_tma_kernel<T, true><<<grid, band::TTHREADS, band::tri_smem_bytes<T, true>(), st>>>(tmV, tmW, p, Bx, By, W, alpha, beta);
      h->launches++;
      SCIMLOP_CUDA(h, cudaGetLastError());
      return SCIMLOP_OK;
    }
  }
  if (bx == 1 && by == 1 && mi % 2 == 0 && aligned16(V) && aligned16(W) && k <= 65535) {
    // 5-point stencil fast path: sliding window over o, 256 i per block
    const int ochunk = 128;
    dim3 g((unsigned)((mi / 2 + 127) / 128), (unsigned)((mo + ochunk - 1) / ochunk), (unsigned)k);
    if (g.y <= 65535u) {
      band::band_kronsum_tri_kernel<T><<<g, 128, 0, st>>>(mi, mo, ochunk, Bx, By, V, W, alpha, beta);
      h->launches++;
      SCIMLOP_CUDA(h, cudaGetLastError());
      return SCIMLOP_OK;
    }
  }
  dim3 grid((unsigned)((mi + 63) / 64), (unsigned)((mo + 3) / 4), (unsigned)std::min<int64_t>(k, 64));
  if (grid.y > 65535u) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kronsum_banded: outer extent too large");
  band::band_kronsum_kernel<T><<<grid, band::THREADS, 0, st>>>(mi, mo, k, bx, by, Bx, By, V, W, alpha, beta);
  h->launches++;
  SCIMLOP_CUDA(h, cudaGetLastError());
  return SCIMLOP_OK;
}

extern "C" int scimlop_kronsum_banded_mul(scimlop_handle_t h, int dtype, const void* Bx, int bx, int64_t mo,
                                          const void* By, int by, int64_t mi, const void* V, int64_t ldv, void* W,
                                          int64_t ldw, int64_t k, const void* alpha, const void* beta, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (int rc = check_dtype(h, dtype)) return rc;
  if (mo < 0 || mi < 0 || k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kronsum_banded: negative size");
  if (bx < 0 || by < 0 || bx > band::MAXB || by > band::MAXB)
    return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "kronsum_banded: half-bandwidth must be in 0..%d", band::MAXB);
  if (mo == 0 || mi == 0 || k == 0) return SCIMLOP_OK;
  if (!Bx || !By || !V || !W) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kronsum_banded: null operand");
  if (ldv != mi * mo || ldw != mi * mo)
    return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kronsum_banded: V and W must be contiguous (ld == mi*mo)");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SCIMLOP_F64)
    return kronsum_banded_t<double>(h, (const double*)Bx, bx, mo, (const double*)By, by, mi, (const double*)V, (double*)W,
                                    k, scalar_or<double>(alpha, 1.0), scalar_or<double>(beta, 0.0), st);
  return kronsum_banded_t<float>(h, (const float*)Bx, bx, mo, (const float*)By, by, mi, (const float*)V, (float*)W, k,
                                 scalar_or<float>(alpha, 1.f), scalar_or<float>(beta, 0.f), st);
}

// ================================================================================================
// operator trees
// ================================================================================================
namespace {

inline bool is_elementwise(int kind) { return kind == SCIMLOP_IDENTITY || kind == SCIMLOP_DIAG || kind == SCIMLOP_BDIAG; }

struct TreeInfo {
  long long rows, cols;
  long long max_inter;  // largest chain intermediate (rows), 0 if none
  int nbuf;
};

int tree_analyse(scimlop_context* c, const scimlop_term_t* terms, int nterms, TreeInfo* ti) {
  if (nterms < 0 || (nterms > 0 && !terms)) return scimlop_fail(c, SCIMLOP_ERR_NULL_POINTER, "tree: terms null");
  ti->rows = -1; ti->cols = -1; ti->max_inter = 0; ti->nbuf = 0;
  for (int t = 0; t < nterms; t++) {
    const scimlop_term_t& T = terms[t];
    if (T.nfactors < 1 || !T.factors) return scimlop_fail(c, SCIMLOP_ERR_BAD_SHAPE, "tree: term %d has no factors", t);
    long long rows = T.factors[0].rows, cols = T.factors[T.nfactors - 1].cols;
    int ndense = 0;
    for (int f = 0; f < T.nfactors; f++) {
      const scimlop_factor_t& F = T.factors[f];
      if (F.kind < SCIMLOP_DENSE || F.kind > SCIMLOP_NULL) return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "tree: factor kind %d", F.kind);
      if (F.rows < 0 || F.cols < 0) return scimlop_fail(c, SCIMLOP_ERR_BAD_SHAPE, "tree: negative factor size");
      if (is_elementwise(F.kind) && F.rows != F.cols) return scimlop_fail(c, SCIMLOP_ERR_BAD_SHAPE, "tree: elementwise factor must be square");
      if (f + 1 < T.nfactors && F.cols != T.factors[f + 1].rows)
        return scimlop_fail(c, SCIMLOP_ERR_BAD_SHAPE, "tree: term %d: factor %d cols (%lld) != factor %d rows (%lld)", t, f,
                            (long long)F.cols, f + 1, (long long)T.factors[f + 1].rows);
      if (F.kind == SCIMLOP_DENSE) ndense++;
    }
    if (ti->rows < 0) { ti->rows = rows; ti->cols = cols; }
    else if (ti->rows != rows || ti->cols != cols)
      return scimlop_fail(c, SCIMLOP_ERR_BAD_SHAPE, "tree: summands have different sizes");
    if (ndense > 0) {
      // intermediates of the chain: the right-hand diagonal run applied to V (if any) and the output of
      // every dense factor except the leftmost one; they ping-pong between at most two buffers
      int ninter = 0;
      int last_dense = -1, first_dense = -1;
      for (int f = 0; f < T.nfactors; f++)
        if (T.factors[f].kind == SCIMLOP_DENSE) { if (first_dense < 0) first_dense = f; last_dense = f; }
      for (int f = last_dense + 1; f < T.nfactors; f++)
        if (T.factors[f].kind == SCIMLOP_DIAG || T.factors[f].kind == SCIMLOP_BDIAG) {
          ninter++; ti->max_inter = std::max<long long>(ti->max_inter, T.factors[last_dense].cols); break;
        }
      for (int f = first_dense + 1; f <= last_dense; f++)
        if (T.factors[f].kind == SCIMLOP_DENSE) { ninter++; ti->max_inter = std::max<long long>(ti->max_inter, T.factors[f].rows); }
      ti->nbuf = std::max(ti->nbuf, std::min(ninter, 2));
    }
  }
  if (ti->rows < 0) { ti->rows = 0; ti->cols = 0; }
  return SCIMLOP_OK;
}

template <typename T>
void set_factor(elt::Factor<T>& dst, const scimlop_factor_t& F) {
  dst.ptr = static_cast<const T*>(F.data);
  if (F.kind == SCIMLOP_BDIAG) { dst.ld = F.ld; dst.div = 1; dst.mod = 1; }
  else { dst.ld = 0; dst.div = 1; dst.mod = F.rows; }
}

// add the elementwise factors [first, last) of a term to (nvec-sorted) slot t of an eltwise program
template <typename T>
int add_elt_factors(scimlop_context* c, elt::Params<T>& p, int t, const scimlop_factor_t* f, int first, int last) {
  int nv = 0, nb = 0;
  const scimlop_factor_t* vec[elt::MAX_FAC];
  const scimlop_factor_t* bat[elt::MAX_FAC];
  for (int i = first; i < last; i++) {
    if (f[i].kind == SCIMLOP_IDENTITY) continue;
    if (nv + nb == elt::MAX_FAC) return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "tree: more than %d diagonal factors in one product", elt::MAX_FAC);
    if (!f[i].data) return scimlop_fail(c, SCIMLOP_ERR_NULL_POINTER, "tree: diagonal factor without data");
    if (f[i].kind == SCIMLOP_DIAG) vec[nv++] = &f[i]; else bat[nb++] = &f[i];
  }
  for (int i = 0; i < nv; i++) set_factor(p.fac[t][i], *vec[i]);
  for (int i = 0; i < nb; i++) set_factor(p.fac[t][nv + i], *bat[i]);
  p.nvec[t] = nv; p.nfac[t] = nv + nb;
  return SCIMLOP_OK;
}

template <typename T>
int tree_mul_t(scimlop_context* c, int precision, const scimlop_term_t* terms, int nterms, const T* V, long long ldv,
               T* W, long long ldw, long long k, T alpha, T beta, void* workspace, size_t ws_bytes, const TreeInfo& ti,
               cudaStream_t st) {
  const long long rows = ti.rows;
  const size_t buf_bytes = align_up((size_t)ti.max_inter * (size_t)k * sizeof(T), 256);
  if (ti.nbuf > 0 && (!workspace || ws_bytes < buf_bytes * ti.nbuf))
    return scimlop_fail(c, SCIMLOP_ERR_WORKSPACE, "tree: workspace %zu B < required %zu B (call cache_operator first)",
                        ws_bytes, buf_bytes * ti.nbuf);
  T* bufs[2] = {static_cast<T*>(workspace), reinterpret_cast<T*>(static_cast<char*>(workspace) + buf_bytes)};

  // ---- pass 1: all elementwise terms in one sweep (also carries beta * W) ----
  bool w_started = false;  // has W already received beta*W_in ?
  {
    elt::Params<T> p;
    memset(&p, 0, sizeof(p));
    p.n = rows; p.k = k; p.V = V; p.ldv = ldv; p.W = W; p.ldw = ldw; p.alpha = alpha; p.beta = beta;
    int ne = 0, ndense_terms = 0;
    for (int t = 0; t < nterms; t++) {
      const scimlop_term_t& Tm = terms[t];
      bool dense = false, null = false;
      for (int f = 0; f < Tm.nfactors; f++) {
        dense = dense || Tm.factors[f].kind == SCIMLOP_DENSE;
        null = null || Tm.factors[f].kind == SCIMLOP_NULL;
      }
      if (null || Tm.scale == 0.0) continue;  // NullOperator / iszero(λ) short-circuit (src/basic.jl:521-523)
      if (dense) { ndense_terms++; continue; }
      if (ne == elt::MAX_TERMS) {
        int rc = launch_eltwise(c, p, st);
        if (rc) return rc;
        w_started = true;
        p.beta = T(1); ne = 0; p.nterms = 0;
      }
      p.coef[ne] = (T)Tm.scale;
      if (int rc = add_elt_factors(c, p, ne, Tm.factors, 0, Tm.nfactors)) return rc;
      ne++;
      p.nterms = ne;
    }
    if (ne > 0 || ndense_terms == 0) {
      // ndense_terms == 0 && ne == 0: the whole tree is null -> W = beta * W with a strong zero
      if (!(w_started && ne == 0)) {
        int rc = launch_eltwise(c, p, st);
        if (rc) return rc;
        w_started = true;
      }
    }
  }

  // ---- pass 2: one fused-epilogue GEMM chain per dense term ----
  for (int t = 0; t < nterms; t++) {
    const scimlop_term_t& Tm = terms[t];
    bool dense = false, null = false;
    int first_dense = -1;
    for (int f = 0; f < Tm.nfactors; f++) {
      if (Tm.factors[f].kind == SCIMLOP_DENSE && first_dense < 0) first_dense = f;
      dense = dense || Tm.factors[f].kind == SCIMLOP_DENSE;
      null = null || Tm.factors[f].kind == SCIMLOP_NULL;
    }
    if (!dense || null || Tm.scale == 0.0) continue;
    const T* cur = V;
    long long cur_ld = ldv;
    int flip = 0;
    int last_dense = first_dense;
    for (int f = first_dense; f < Tm.nfactors; f++) if (Tm.factors[f].kind == SCIMLOP_DENSE) last_dense = f;
    // diagonal factors right of the rightmost dense factor act on V: one elementwise pass into scratch
    {
      bool any = false;
      for (int f = last_dense + 1; f < Tm.nfactors; f++)
        any = any || Tm.factors[f].kind == SCIMLOP_DIAG || Tm.factors[f].kind == SCIMLOP_BDIAG;
      if (any) {
        elt::Params<T> p;
        memset(&p, 0, sizeof(p));
        const long long in_rows = Tm.factors[last_dense].cols;
        T* tmp = bufs[flip]; flip ^= 1;
        p.n = in_rows; p.k = k; p.V = cur; p.ldv = cur_ld; p.W = tmp; p.ldw = in_rows; p.alpha = T(1); p.beta = T(0);
        p.nterms = 1; p.coef[0] = T(1);
        if (int rc = add_elt_factors(c, p, 0, Tm.factors, last_dense + 1, Tm.nfactors)) return rc;
        if (int rc = launch_eltwise(c, p, st)) return rc;
        cur = tmp; cur_ld = in_rows;
      }
    }
    for (int f = last_dense; f >= first_dense; f--) {
      const scimlop_factor_t& F = Tm.factors[f];
      if (F.kind != SCIMLOP_DENSE) continue;
      if (!F.data) return scimlop_fail(c, SCIMLOP_ERR_NULL_POINTER, "tree: dense factor without data");
      const long long in_rows = F.cols, out_rows = F.rows;
      if (F.ld < std::max<long long>(1, F.trans ? F.cols : F.rows))
        return scimlop_fail(c, SCIMLOP_ERR_BAD_SHAPE, "tree: dense factor ld too small");
      // the diagonal factors between this dense factor and the next one to its left scale the rows of its
      // output: they ride in the GEMM epilogue
      ScaleArg sc[dmma::MAX_SCALES];
      int ns = 0;
      for (int q = f - 1; q >= 0 && Tm.factors[q].kind != SCIMLOP_DENSE; q--) {
        const scimlop_factor_t& D = Tm.factors[q];
        if (D.kind == SCIMLOP_IDENTITY) continue;
        if (ns == dmma::MAX_SCALES)
          return scimlop_fail(c, SCIMLOP_ERR_UNSUPPORTED, "tree: more than %d diagonal factors left of a dense factor", dmma::MAX_SCALES);
        if (!D.data) return scimlop_fail(c, SCIMLOP_ERR_NULL_POINTER, "tree: diagonal factor without data");
        sc[ns].ptr = D.data; sc[ns].ld = (D.kind == SCIMLOP_BDIAG) ? D.ld : 0; ns++;
      }
      if (f == first_dense) {
        // leftmost dense factor: writes W with λ, α and β fused as well
        const T b2 = w_started ? T(1) : beta;
        int rc = gemm_dispatch<T>(c, precision, F.trans != 0, false, out_rows, k, in_rows, (T)(alpha * (T)Tm.scale),
                                  static_cast<const T*>(F.data), F.ld, 0, cur, cur_ld, 0, b2, W, ldw, 0, 1, ns, sc, st);
        if (rc) return rc;
        w_started = true;
      } else {
        T* out = bufs[flip]; flip ^= 1;
        int rc = gemm_dispatch<T>(c, precision, F.trans != 0, false, out_rows, k, in_rows, T(1),
                                  static_cast<const T*>(F.data), F.ld, 0, cur, cur_ld, 0, T(0), out, out_rows, 0, 1, ns,
                                  sc, st);
        if (rc) return rc;
        cur = out; cur_ld = out_rows;
      }
    }
  }
  return SCIMLOP_OK;
}

}  // namespace

extern "C" int scimlop_tree_workspace_bytes(int dtype, const scimlop_term_t* terms, int nterms, int64_t k,
                                            size_t* bytes) {
  if (!bytes) return SCIMLOP_ERR_NULL_POINTER;
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32) return SCIMLOP_ERR_UNSUPPORTED;
  if (k < 0) return SCIMLOP_ERR_BAD_SHAPE;
  TreeInfo ti;
  if (int rc = tree_analyse(nullptr, terms, nterms, &ti)) return rc;
  *bytes = ti.nbuf * align_up((size_t)ti.max_inter * (size_t)k * elem_size(dtype), 256);
  return SCIMLOP_OK;
}

extern "C" int scimlop_tree_mul(scimlop_handle_t h, int dtype, int precision, const scimlop_term_t* terms, int nterms,
                                const void* V, int64_t ldv, void* W, int64_t ldw, int64_t k, const void* alpha,
                                const void* beta, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (int rc = check_dtype(h, dtype)) return rc;
  if (k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "tree: negative k");
  TreeInfo ti;
  if (int rc = tree_analyse(h, terms, nterms, &ti)) return rc;
  if (nterms == 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "tree: no terms (use scimlop_axpby for a pure beta*W)");
  if (k == 0 || ti.rows == 0) return SCIMLOP_OK;
  if (!W || !V) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "tree: V/W null");
  if (ldw < ti.rows || ldv < ti.cols) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "tree: leading dimension smaller than rows");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SCIMLOP_F64)
    return tree_mul_t<double>(h, precision, terms, nterms, (const double*)V, ldv, (double*)W, ldw, k,
                              scalar_or<double>(alpha, 1.0), scalar_or<double>(beta, 0.0), workspace, workspace_bytes, ti, st);
  return tree_mul_t<float>(h, precision, terms, nterms, (const float*)V, ldv, (float*)W, ldw, k,
                           scalar_or<float>(alpha, 1.f), scalar_or<float>(beta, 0.f), workspace, workspace_bytes, ti, st);
}

extern "C" int scimlop_shard_columns(int64_t k, int nranks, int rank, int64_t* first, int64_t* count) {
  if (!first || !count) return SCIMLOP_ERR_NULL_POINTER;
  if (k < 0 || nranks < 1 || rank < 0 || rank >= nranks) return SCIMLOP_ERR_BAD_SHAPE;
  const int64_t base = k / nranks, rem = k % nranks;
  *first = rank * base + std::min<int64_t>(rank, rem);
  *count = base + (rank < rem ? 1 : 0);
  return SCIMLOP_OK;
}
