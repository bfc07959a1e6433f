// lu.cu -- the factorised solve behind `ldiv!`, Float64.
//
// The reference caches an LU (`factorize(L::MatrixOperator)` -> `lu`, src/matrix.jl:513-516; per Kronecker factor,
// src/tensor.jl:227) and solves with it (`ldiv!`: src/tensor.jl:612-673, outer_div! :864-899 -- LAPACK getrs around two
// permutedims!).  Round 1 cached an explicit inverse from an unblocked Gauss-Jordan (2n+1 launches, 263 ms at 4096^2,
// residual ~ cond(A) * eps).  This file is the replacement:
//
//   factorise  blocked right-looking LU with partial pivoting, P A = L U.  Panels of 32 columns are factorised by ONE
//              multi-CTA kernel per panel: every thread owns one row of the panel in registers, a column step is a
//              block arg-max, a grid barrier, the exchange of the pivot row through global memory, a second barrier and
//              32 FMAs per thread.  Row swaps + the 32 x w triangular solve of the block row are one more kernel; the
//              trailing update is the DMMA GEMM (alpha = -1, beta = 1).  With LOOK-AHEAD: the panel kernel (16 CTAs,
//              ~3.3 us of barriers per column) is the critical path, so only the next panel's 32 columns are brought up
//              to date on the caller's stream (one small kernel for their swaps + block-row solve; their rank-32 update
//              rides in the next panel kernel's load phase), while the swaps, block-row solve and trailing update of
//              every other column run on the handle's side stream on the remaining SMs.  Measured 2.5 / 4.9 / 10.3 /
//              23.7 ms at n = 512 / 1024 / 2048 / 4096 including the solve structures, the explicit inverse and
//              cond_1 (sequential: 2.9 / 5.7 / 12.2 / 30.1; round 1: 4.4 / 10.8 / 30 / 263); timeline in
//              profiles/r02_lu_trace.txt (tools/lu_trace.py): 137 us per panel step = 113 panel + 18 prep + gaps.
//   solve      triangular solves as GEMMs: 128 x 128 diagonal blocks of L and U are inverted once (in shared memory),
//              the off-diagonal block rows are pre-multiplied by them (L' = Dinv L, U' = Dinv U), and a block step is
//              x_i = Dinv_ii b_i  followed by  x_i -= L'_{i,<i} x_{<i}  -- both DMMA GEMMs, out of place between two
//              buffers (no in-place hazards for any GEMM kernel the dispatcher may choose), acting along ANY Kronecker
//              mode exactly like the mode products of mul! (no permutedims).  The row permutation is one gather pass.
//              Cost: 1.25x the flops of a multiplication by the explicit inverse at n = 512 (block triangles), with the
//              backward stability of substitution: residual ~ eps * |A| |x| (inverted diagonal blocks only, the scheme of
//              cuBLAS / MAGMA trsm).
//   inverse    the explicit inverse is ALSO formed (the solve applied to the identity: GEMM speed) together with
//              cond_1(A) = |A|_1 |Ainv|_1, returned to the caller: for a well-conditioned factor the host keeps the
//              round-1 fast path (ldiv! = mul! with the inverse factors, residual cond * eps still inside the 1e-12
//              budget), for an ill-conditioned one it takes the factorised solve.
#include <cstdio>
#include <algorithm>
#include <cstring>

#include "context.h"
#include "ptx.cuh"

namespace {

using scimlop::cluster_sync_all;
using scimlop::ld_cluster;
using scimlop::map_to_rank;
using scimlop::smem_u32;

constexpr int PB = 32;        // panel width
constexpr int PROWS = 256;    // panel rows per CTA (one per thread)
constexpr int DB = 128;       // diagonal block of the triangular solves (= DMMA tile rows)

inline long long ld_of(long long n) { return (n + 3) / 4 * 4; }
inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// ---- blob layout (device memory, caller-owned) --------------------------------------------------------------------
struct Blob {
  int* perm;      // n: row r of P*b is row perm[r] of b
  double* LU;     // ldf x n: the LAPACK-style factors (unit lower + upper), pivoted
  double* LUr;    // ldf x n: off-diagonal blocks pre-multiplied by the inverse diagonal block of their block ROW
  double* DL;     // ldf x DB: inverse diagonal blocks of L, block i at rows i*DB
  double* DU;     // ldf x DB: inverse diagonal blocks of U
  long long ldf;
};
inline size_t blob_bytes(long long n) {
  const long long ldf = ld_of(n);
  return align256((size_t)n * 4) + 2 * align256((size_t)ldf * n * 8) + 2 * align256((size_t)ldf * DB * 8);
}
inline Blob blob_at(void* base, long long n) {
  Blob b;
  b.ldf = ld_of(n);
  char* p = static_cast<char*>(base);
  b.perm = reinterpret_cast<int*>(p); p += align256((size_t)n * 4);
  b.LU = reinterpret_cast<double*>(p); p += align256((size_t)b.ldf * n * 8);
  b.LUr = reinterpret_cast<double*>(p); p += align256((size_t)b.ldf * n * 8);
  b.DL = reinterpret_cast<double*>(p); p += align256((size_t)b.ldf * DB * 8);
  b.DU = reinterpret_cast<double*>(p);
  return b;
}
// scratch: piv (n ints), sync words, pivot-row exchange, norms / info, two n x n work matrices
struct Scratch {
  int* piv;
  unsigned* bar;        // one counter per panel
  double* slots;        // [2 parities][max CTAs][PB + 2] candidate rows, then [2][PB] diagonal rows
  double* norms;        // [0] = |A|_1, [1] = |Ainv|_1 (as doubles; max via atomics on the bit pattern)
  int* info;
  double* T1;
  double* T2;
};
inline size_t scratch_bytes(long long n) {
  const long long ldf = ld_of(n);
  const long long npan = (n + PB - 1) / PB, ncta = (n + PROWS - 1) / PROWS;
  return align256((size_t)n * 4) + align256((size_t)npan * 4) + align256((size_t)(2 * ncta * (PB + 2) + 2 * PB) * 8) +
         256 + 256 + 2 * align256((size_t)ldf * n * 8);
}
inline Scratch scratch_at(void* base, long long n) {
  const long long ldf = ld_of(n);
  const long long npan = (n + PB - 1) / PB, ncta = (n + PROWS - 1) / PROWS;
  Scratch s;
  char* p = static_cast<char*>(base);
  s.piv = reinterpret_cast<int*>(p); p += align256((size_t)n * 4);
  s.bar = reinterpret_cast<unsigned*>(p); p += align256((size_t)npan * 4);
  s.slots = reinterpret_cast<double*>(p); p += align256((size_t)(2 * ncta * (PB + 2) + 2 * PB) * 8);
  s.norms = reinterpret_cast<double*>(p); p += 256;
  s.info = reinterpret_cast<int*>(p); p += 256;
  s.T1 = reinterpret_cast<double*>(p); p += align256((size_t)ldf * n * 8);
  s.T2 = reinterpret_cast<double*>(p);
  return s;
}

// ---- kernels -----------------------------------------------------------------------------------------------------------
__global__ void copy_matrix_kernel(long long n, const double* __restrict__ A, long long lda, double* __restrict__ B, long long ldb) {
  const long long total = n * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    B[(i % n) + (i / n) * ldb] = A[(i % n) + (i / n) * lda];
}

// max column abs-sum of an n x n matrix -> out (non-negative doubles compare like their bit patterns)
__global__ void norm1_kernel(long long n, const double* __restrict__ A, long long lda, double* out) {
  __shared__ double part[32];
  for (long long c = blockIdx.x; c < n; c += gridDim.x) {
    double s = 0.0;
    for (long long r = threadIdx.x; r < n; r += blockDim.x) s += fabs(A[r + c * lda]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      s = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(s));
    }
    __syncthreads();
  }
}

// a[c] of a register array with a run-time c: a jump table (BRX), a handful of executed instructions.  (A chain of 32
// selects costs ~100 per access; unrolling the whole column loop instead made the panel kernels instruction-cache bound:
// "no_instructions" 46 % of the stalls, profiles/r02_lu_panel_ncu_summary.txt.)
#define LU_CASE(k) case k: return a[k];
__device__ __forceinline__ double reg_get(const double (&a)[PB], int c) {
  switch (c) {
    LU_CASE(0) LU_CASE(1) LU_CASE(2) LU_CASE(3) LU_CASE(4) LU_CASE(5) LU_CASE(6) LU_CASE(7) LU_CASE(8) LU_CASE(9) LU_CASE(10)
    LU_CASE(11) LU_CASE(12) LU_CASE(13) LU_CASE(14) LU_CASE(15) LU_CASE(16) LU_CASE(17) LU_CASE(18) LU_CASE(19) LU_CASE(20)
    LU_CASE(21) LU_CASE(22) LU_CASE(23) LU_CASE(24) LU_CASE(25) LU_CASE(26) LU_CASE(27) LU_CASE(28) LU_CASE(29) LU_CASE(30)
    default: return a[31];
  }
}
#undef LU_CASE

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned v;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

// Look-ahead, fused into the panel kernels: the rows of this panel lie below the previous panel (columns [jp, jp + wp)),
// whose rank-wp update of THIS panel's columns is applied to the register copy right after the load:
// a(c) -= sum_k L(r, jp + k) U(jp + k, j0 + c), U (already swapped and solved by lu_lookahead_prep_kernel) staged in
// shared memory.  The panel kernel stores its columns at the end, so the un-updated values in memory are simply replaced.
__device__ __forceinline__ void panel_apply_previous(double (&a)[PB], const double* __restrict__ A, long long lda, long long r, bool live,
                                                     int j0, int w, int jp, int wp, double* Us /* PB * PB */) {
  if (wp <= 0) return;                              // uniform over the grid
  for (int q = threadIdx.x; q < PB * PB; q += blockDim.x) {
    const int k = q % PB, cc = q / PB;
    Us[q] = (k < wp && cc < w) ? A[(long long)(jp + k) + (long long)(j0 + cc) * lda] : 0.0;
  }
  double l[PB];
#pragma unroll
  for (int k = 0; k < PB; k++) l[k] = (live && k < wp) ? A[r + (long long)(jp + k) * lda] : 0.0;
  __syncthreads();
#pragma unroll
  for (int c = 0; c < PB; c++) {
    double acc = a[c];
#pragma unroll
    for (int k = 0; k < PB; k++) acc -= l[k] * Us[k + c * PB];
    a[c] = acc;
  }
  __syncthreads();                                  // Us may alias the kernel's other shared arrays later: done with it
}

// One panel: columns j0 .. j0 + w - 1, rows j0 .. n - 1.  Thread t of CTA b owns row j0 + PROWS * b + t: its w panel entries
// live in registers for the whole panel.  All CTAs are co-resident (<= n / 256 of them), so the software barrier is safe.
// A column step has ONE grid barrier: before it every CTA publishes its candidate pivot (|value|, row) TOGETHER with that
// row's panel entries, and CTA 0 also publishes the diagonal row; after it every CTA picks the same winner and reads the
// winner's row from the winner's slot.  Slots are double-buffered by column parity (a CTA runs at most one column ahead).
__global__ void __launch_bounds__(PROWS) lu_panel_kernel(long long n, double* __restrict__ A, long long lda, int j0, int w,
                                                         int* __restrict__ piv, unsigned* bar, double* slots, int* info, int jp, int wp) {
  __shared__ double Us[PB * PB];
  // slots: [parity][CTA][PB + 2]: candidate row values, |value|, row index (as a double);  diag: [parity][PB] after them
  __shared__ double sval[PROWS / 32];
  __shared__ int srow[PROWS / 32];
  __shared__ int swin, spiv;
  constexpr int SL = PB + 2;
  const int nct = gridDim.x;
  double* diag = slots + 2 * (size_t)nct * SL;
  const int i = blockIdx.x * PROWS + threadIdx.x;     // panel-relative row
  const long long r = (long long)j0 + i;
  const bool live = r < n;
  double a[PB];
#pragma unroll
  for (int c = 0; c < PB; c++) a[c] = (live && c < w) ? A[r + (long long)(j0 + c) * lda] : 0.0;
  panel_apply_previous(a, A, lda, r, live, j0, w, jp, wp, Us);
  unsigned epoch = 0;
  for (int c = 0; c < w; c++) {
    const int par = c & 1;
    // ---- local arg-max of |a(:, c)| over rows >= c (first maximum, like idamax) ----
    const double ac = reg_get(a, c);
    double best = (live && i >= c) ? fabs(ac) : -1.0;
    if (best != best) best = 1.0e308;                 // NaN: take it as the pivot so that it propagates like LAPACK's
    int bi = i;
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, best, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sval[threadIdx.x >> 5] = best; srow[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int q = 1; q < PROWS / 32; q++)
        if (sval[q] > best || (sval[q] == best && srow[q] < bi)) { best = sval[q]; bi = srow[q]; }
      swin = bi;
      double* sl = slots + ((size_t)par * nct + blockIdx.x) * SL;
      sl[PB] = best;
      sl[PB + 1] = (double)bi;
    }
    __syncthreads();
    if (i == swin) {                                   // this CTA's candidate row publishes its entries
      double* sl = slots + ((size_t)par * nct + blockIdx.x) * SL;
#pragma unroll
      for (int q = 0; q < PB; q++) sl[q] = a[q];
    }
    if (i == c) {                                      // the diagonal row (always in CTA 0)
#pragma unroll
      for (int q = 0; q < PB; q++) diag[par * PB + q] = a[q];
    }
    grid_barrier(bar, (++epoch) * nct);
    // ---- global arg-max (every CTA computes the same answer): warp 0 reads all slots at once ----
    if (threadIdx.x < 32) {
      double gb = -1.0;
      int gi = 0x7fffffff, gw = 0;
      for (int q = threadIdx.x; q < nct; q += 32) {
        const double* sl = slots + ((size_t)par * nct + q) * SL;
        const double v = __ldcg(sl + PB);
        const int ri = (int)__ldcg(sl + PB + 1);
        if (v > gb || (v == gb && ri < gi)) { gb = v; gi = ri; gw = q; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, gb, o);
        const int oi = __shfl_down_sync(0xffffffffu, gi, o);
        const int ow = __shfl_down_sync(0xffffffffu, gw, o);
        if (ov > gb || (ov == gb && oi < gi)) { gb = ov; gi = oi; gw = ow; }
      }
      if (threadIdx.x == 0) {
        spiv = gi;
        swin = gw;
        if (blockIdx.x == 0) {
          piv[j0 + c] = j0 + gi;
          if (!(gb > 0.0)) atomicCAS(info, 0, j0 + c + 1);
        }
      }
    }
    __syncthreads();
    const int p = spiv;
    const double* win = slots + ((size_t)par * nct + swin) * SL;
    double pr[PB];
#pragma unroll
    for (int q = 0; q < PB; q++) pr[q] = __ldcg(win + q);
    if (live && p != c) {
      if (i == c) {
#pragma unroll
        for (int q = 0; q < PB; q++) a[q] = pr[q];
      } else if (i == p) {
#pragma unroll
        for (int q = 0; q < PB; q++) a[q] = __ldcg(diag + par * PB + q);
      }
    }
    // ---- eliminate: l = a(i, c) / pivot, a(i, c+1:) -= l * pivot_row(c+1:) ----
    const double pv = reg_get(pr, c);
    if (live && i > c && pv != 0.0) {
      const double l = reg_get(a, c) / pv;
#pragma unroll
      for (int q = 0; q < PB; q++) {
        if (q == c) a[q] = l;
        if (q > c) a[q] -= l * pr[q];
      }
    }
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < PB; c++) if (c < w) A[r + (long long)(j0 + c) * lda] = a[c];
  }
}

// The same panel step with the CTAs of ONE thread-block cluster (panels of up to 16 x 256 = 4096 rows): candidates and
// the diagonal row live in each CTA's shared memory and are read through distributed shared memory after one cluster
// barrier per column -- ~0.3 us round trips instead of the ~1 us L2 round trips of the grid-barrier form above.
__global__ void __launch_bounds__(PROWS) lu_panel_cluster_kernel(long long n, double* __restrict__ A, long long lda, int j0, int w,
                                                                 int* __restrict__ piv, int* info, int jp, int wp) {
  __shared__ double Us[PB * PB];
  __shared__ double slot[2][PB + 2];     // this CTA's candidate row: entries, |value|, row index
  __shared__ double diag[2][PB];         // the diagonal row (rank 0's copy is the one that is read)
  __shared__ double prow[PB], drow[PB];
  __shared__ double sval[PROWS / 32];
  __shared__ int srow[PROWS / 32];
  __shared__ int swin, spiv;
  const int nct = gridDim.x;             // = cluster size
  const int i = blockIdx.x * PROWS + threadIdx.x;
  const long long r = (long long)j0 + i;
  const bool live = r < n;
  double a[PB];
#pragma unroll
  for (int c = 0; c < PB; c++) a[c] = (live && c < w) ? A[r + (long long)(j0 + c) * lda] : 0.0;
  panel_apply_previous(a, A, lda, r, live, j0, w, jp, wp, Us);
  for (int c = 0; c < w; c++) {
    const int par = c & 1;
    const double ac = reg_get(a, c);
    double best = (live && i >= c) ? fabs(ac) : -1.0;
    if (best != best) best = 1.0e308;
    int bi = i;
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, best, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sval[threadIdx.x >> 5] = best; srow[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int q = 1; q < PROWS / 32; q++)
        if (sval[q] > best || (sval[q] == best && srow[q] < bi)) { best = sval[q]; bi = srow[q]; }
      swin = bi;
      slot[par][PB] = best;
      slot[par][PB + 1] = (double)bi;
    }
    __syncthreads();
    if (i == swin) {
#pragma unroll
      for (int q = 0; q < PB; q++) slot[par][q] = a[q];
    }
    if (i == c) {
#pragma unroll
      for (int q = 0; q < PB; q++) diag[par][q] = a[q];
    }
    cluster_sync_all();
    if (threadIdx.x < 32) {
      double gb = -1.0;
      int gi = 0x7fffffff, gw = 0;
      for (int q = threadIdx.x; q < nct; q += 32) {
        const double v = ld_cluster<double>(map_to_rank(smem_u32(&slot[par][PB]), (uint32_t)q));
        const int ri = (int)ld_cluster<double>(map_to_rank(smem_u32(&slot[par][PB + 1]), (uint32_t)q));
        if (v > gb || (v == gb && ri < gi)) { gb = v; gi = ri; gw = q; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, gb, o);
        const int oi = __shfl_down_sync(0xffffffffu, gi, o);
        const int ow = __shfl_down_sync(0xffffffffu, gw, o);
        if (ov > gb || (ov == gb && oi < gi)) { gb = ov; gi = oi; gw = ow; }
      }
      gw = __shfl_sync(0xffffffffu, gw, 0);
      gi = __shfl_sync(0xffffffffu, gi, 0);
      gb = __shfl_sync(0xffffffffu, gb, 0);
      // lane q fetches entry q of the winner's row and of the diagonal row into local shared memory
      prow[threadIdx.x] = ld_cluster<double>(map_to_rank(smem_u32(&slot[par][threadIdx.x]), (uint32_t)gw));
      drow[threadIdx.x] = ld_cluster<double>(map_to_rank(smem_u32(&diag[par][threadIdx.x]), 0u));
      if (threadIdx.x == 0) {
        spiv = gi;
        if (blockIdx.x == 0) {
          piv[j0 + c] = j0 + gi;
          if (!(gb > 0.0)) atomicCAS(info, 0, j0 + c + 1);
        }
      }
    }
    __syncthreads();
    const int p = spiv;
    double pr[PB];
#pragma unroll
    for (int q = 0; q < PB; q++) pr[q] = prow[q];
    if (live && p != c) {
      if (i == c) {
#pragma unroll
        for (int q = 0; q < PB; q++) a[q] = pr[q];
      } else if (i == p) {
#pragma unroll
        for (int q = 0; q < PB; q++) a[q] = drow[q];
      }
    }
    const double pv = reg_get(pr, c);
    if (live && i > c && pv != 0.0) {
      const double l = reg_get(a, c) / pv;
#pragma unroll
      for (int q = 0; q < PB; q++) {
        if (q == c) a[q] = l;
        if (q > c) a[q] -= l * pr[q];
      }
    }
    __syncthreads();     // prow / drow are rewritten by warp 0 in the next column
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < PB; c++) if (c < w) A[r + (long long)(j0 + c) * lda] = a[c];
  }
  cluster_sync_all();    // no CTA may exit while a peer can still read its shared memory
}

// Row swaps of the panel applied to every column outside it, then (columns right of the panel) the w x w unit-lower
// solve that turns A12 into U12.  One thread per column.  (A variant that loaded all rows up front and emulated the swap
// sequence in shared memory was slower: 45 vs 34 us per panel at n = 4096.)
constexpr int ST_THREADS = 128;
__global__ void __launch_bounds__(ST_THREADS) lu_swap_trsm_kernel(long long n, double* __restrict__ A, long long lda, int j0, int w,
                                                                  const int* __restrict__ piv, long long ncols, long long skip0,
                                                                  long long skipw) {
  __shared__ double L11[PB * PB];
  __shared__ int sp[PB];
  for (int q = threadIdx.x; q < PB * PB; q += blockDim.x) {
    const int rr = q % PB, cc = q / PB;
    L11[q] = (rr < w && cc < w && rr > cc) ? A[(long long)(j0 + rr) + (long long)(j0 + cc) * lda] : 0.0;
  }
  if (threadIdx.x < PB) sp[threadIdx.x] = threadIdx.x < w ? piv[j0 + threadIdx.x] : 0;
  __syncthreads();
  long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  if (col >= skip0) col += skipw;                  // skip the panel's own columns (and, with look-ahead, the next panel's)
  double* a = A + col * lda;
  for (int c = 0; c < w; c++) {
    const int p = sp[c];
    if (p != j0 + c) {
      const double t = a[j0 + c];
      a[j0 + c] = a[p];
      a[p] = t;
    }
  }
  if (col > j0) {
    double x[PB];
#pragma unroll
    for (int q = 0; q < PB; q++) x[q] = q < w ? a[j0 + q] : 0.0;
#pragma unroll
    for (int q = 1; q < PB; q++) {
      double sacc = x[q];
#pragma unroll
      for (int k2 = 0; k2 < PB; k2++) if (k2 < q) sacc -= L11[q + k2 * PB] * x[k2];
      x[q] = sacc;
    }
#pragma unroll
    for (int q = 0; q < PB; q++) if (q < w) a[j0 + q] = x[q];
  }
}

// ---- look-ahead: the NEXT panel's columns, on the critical path ------------------------------------------------------------
// Panel p's row swaps, block-row solve and rank-w update applied to the wn columns [c0, c0 + wn) of panel p + 1 only, so
// that panel p + 1 can start while a second stream does the same for all other columns.  The generic kernels cost
// ~35 us (w dependent swap steps per column) + a GEMM launch here; these two take a few us:
//   prep    (one CTA) the w swaps touch at most 2 w rows: their net permutation is formed once in shared memory, every
//           affected element is loaded (independent loads), then stored; then the w x w unit-lower solve per column.
//   update  rows below the panel: C(r, :) -= L(r, 0..w) U(0..w, :) is applied by the NEXT panel kernel to its register
//           copy of those rows while it loads them (panel_apply_previous): no launch, no extra pass.
constexpr int LA_THREADS = 256;
__global__ void __launch_bounds__(LA_THREADS) lu_lookahead_prep_kernel(double* __restrict__ A, long long lda, int j0, int w, int c0, int wn,
                                                                       const int* __restrict__ piv) {
  __shared__ double L11[PB * PB];
  __shared__ double T[2 * PB][PB + 1];             // the affected rows of the wn columns, after the permutation
  __shared__ int rows[2 * PB], cur[2 * PB];
  __shared__ int nrows_s;
  for (int q = threadIdx.x; q < PB * PB; q += blockDim.x) {
    const int rr = q % PB, cc = q / PB;
    L11[q] = (rr < w && cc < w && rr > cc) ? A[(long long)(j0 + rr) + (long long)(j0 + cc) * lda] : 0.0;
  }
  if (threadIdx.x < 32) {
    // net permutation of the <= 2 w rows the w swaps touch, by one warp: lane i holds rows[w + i] (the pivot rows below
    // the panel met so far) in a register, so "seen before?" is one ballot instead of a scan
    const int lane = threadIdx.x;
    const int myp = lane < w ? piv[j0 + lane] : 0;
    int extra = -1, nr = w;                         // extra: the row id stored at rows[w + lane]
    int mycur = lane;                               // cur[lane] for lane < w; cur[w + lane] kept in xcur
    int xcur = w + lane;
    for (int c = 0; c < w; c++) {
      const int p = __shfl_sync(0xffffffffu, myp, c);
      if (p == j0 + c) continue;
      int b;
      if (p < j0 + w) b = p - j0;
      else {
        const unsigned hit = __ballot_sync(0xffffffffu, extra == p);
        if (hit) b = w + (__ffs(hit) - 1);
        else { b = nr; if (lane == nr - w) extra = p; nr++; }
      }
      // swap cur[c] <-> cur[b]
      const int vc = __shfl_sync(0xffffffffu, mycur, c);
      const int vb = b < w ? __shfl_sync(0xffffffffu, mycur, b) : __shfl_sync(0xffffffffu, xcur, b - w);
      if (lane == c) mycur = vb;
      if (b < w) { if (lane == b) mycur = vc; }
      else if (lane == b - w) xcur = vc;
    }
    if (lane < w) { rows[lane] = j0 + lane; cur[lane] = mycur; }
    if (lane < nr - w) { rows[w + lane] = extra; cur[w + lane] = xcur; }
    if (lane == 0) nrows_s = nr;
  }
  __syncthreads();
  const int nr = nrows_s, total = nr * wn;        // <= 64 * 32 elements: new content of rows[i] = old content of rows[cur[i]]
  double v[2 * PB * PB / LA_THREADS];
#pragma unroll
  for (int u = 0; u < 2 * PB * PB / LA_THREADS; u++) {
    const int idx = threadIdx.x + u * LA_THREADS;
    v[u] = idx < total ? A[(long long)rows[cur[idx % nr]] + (long long)(c0 + idx / nr) * lda] : 0.0;
  }
#pragma unroll
  for (int u = 0; u < 2 * PB * PB / LA_THREADS; u++) {
    const int idx = threadIdx.x + u * LA_THREADS;
    if (idx < total) T[idx % nr][idx / nr] = v[u];
  }
  __syncthreads();                                 // every load of the old content has completed (values are in T)
  if (threadIdx.x < wn) {                          // w x w unit-lower solve on the panel rows of column threadIdx.x
    const int cc = threadIdx.x;
    double x[PB];
#pragma unroll
    for (int q = 0; q < PB; q++) x[q] = q < w ? T[q][cc] : 0.0;
#pragma unroll
    for (int q = 1; q < PB; q++) {
      double sacc = x[q];
#pragma unroll
      for (int k2 = 0; k2 < PB; k2++) if (k2 < q) sacc -= L11[q + k2 * PB] * x[k2];
      x[q] = sacc;
    }
#pragma unroll
    for (int q = 0; q < PB; q++) if (q < w) T[q][cc] = x[q];
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 2 * PB * PB / LA_THREADS; u++) {
    const int idx = threadIdx.x + u * LA_THREADS;
    if (idx < total) A[(long long)rows[idx % nr] + (long long)(c0 + idx / nr) * lda] = T[idx % nr][idx / nr];
  }
}
// perm from the pivot sequence (sequential swaps of the identity), in shared memory by one thread
__global__ void perm_kernel(int n, const int* __restrict__ piv, int* __restrict__ perm) {
  extern __shared__ int sperm[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) sperm[i] = i;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < n; i++) {
      const int p = piv[i];
      if (p != i) { const int t = sperm[i]; sperm[i] = sperm[p]; sperm[p] = t; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = sperm[i];
}

// Inverse of one DB x DB diagonal block, in place in shared memory (LAPACK trti2 order): blocks [0, nb) are the
// unit-lower blocks of L, [nb, 2 nb) the upper blocks of U.  One CTA of DB threads per block.
__global__ void __launch_bounds__(DB) tri_inv_kernel(long long n, const double* __restrict__ LU, long long ldf,
                                                     double* __restrict__ DL, double* __restrict__ DU, int nblocks) {
  extern __shared__ double T[];                    // DB x DB, column-major, pitch DB + 1
  constexpr int P = DB + 1;
  const bool upper = blockIdx.x >= nblocks;
  const int blk = upper ? blockIdx.x - nblocks : blockIdx.x;
  const long long o = (long long)blk * DB;
  const int sz = (int)(n - o < DB ? n - o : DB);
  const int t = threadIdx.x;
  for (int c = 0; c < sz; c++) {
    double v = 0.0;
    if (t < sz) {
      const double x = LU[(o + t) + (o + c) * ldf];
      if (upper) v = t <= c ? x : 0.0;
      else v = t > c ? x : (t == c ? 1.0 : 0.0);
    }
    T[t + c * P] = v;
  }
  __syncthreads();
  if (upper) {
    // columns left to right: T(0:j, j) = -Tinv(0:j, 0:j) * T(0:j, j) / T(j, j)
    for (int j = 0; j < sz; j++) {
      const double d = 1.0 / T[j + j * P];
      double s = 0.0;
      if (t < j) {
        for (int k2 = t; k2 < j; k2++) s += T[t + k2 * P] * T[k2 + j * P];   // Tinv(t, k2) (already inverted) * T(k2, j)
      }
      __syncthreads();
      if (t < j) T[t + j * P] = -s * d;
      if (t == j) T[j + j * P] = d;
      __syncthreads();
    }
  } else {
    // unit lower, columns right to left: T(j+1:, j) = -Tinv(j+1:, j+1:) * T(j+1:, j)
    for (int j = sz - 1; j >= 0; j--) {
      double s = 0.0;
      if (t > j && t < sz) {
        for (int k2 = j + 1; k2 <= t; k2++) s += T[t + k2 * P] * T[k2 + j * P];
      }
      __syncthreads();
      if (t > j && t < sz) T[t + j * P] = -s;
      __syncthreads();
    }
  }
  double* D = upper ? DU : DL;
  for (int c = 0; c < DB; c++)
    if (t < sz) D[(o + t) + (long long)c * ldf] = c < sz ? T[t + c * P] : 0.0;
}

// out(l, r, b) = in(l, perm[r], b) for an (Lp x n x R) view: the row permutation of P*b along one Kronecker mode
__global__ void perm_mode_kernel(long long Lp, long long n, long long R, const int* __restrict__ perm,
                                 const double* __restrict__ in, double* __restrict__ out) {
  const long long total = Lp * n * R;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long l = i % Lp, r = (i / Lp) % n, b = i / (Lp * n);
    out[i] = in[l + Lp * ((long long)perm[r] + n * b)];
  }
}
// out(l, perm[r], b) = in(l, r, b): P^T along one mode (the last pass of a transposed solve)
__global__ void perm_mode_scatter_kernel(long long Lp, long long n, long long R, const int* __restrict__ perm,
                                         const double* __restrict__ in, double* __restrict__ out) {
  const long long total = Lp * n * R;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long l = i % Lp, r = (i / Lp) % n, b = i / (Lp * n);
    out[l + Lp * ((long long)perm[r] + n * b)] = in[i];
  }
}
__global__ void identity_perm_kernel(long long n, long long ldt, const int* __restrict__ perm, double* __restrict__ T) {
  const long long total = n * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i % n, c = i / n;
    T[r + c * ldt] = perm[r] == c ? 1.0 : 0.0;
  }
}

// ---- host drivers --------------------------------------------------------------------------------------------------------
int gemm64(scimlop_context* h, bool ta, bool tb, long long M, long long N, long long K, double alpha, const double* A,
           long long lda, long long sA, const double* B, long long ldb, long long sB, double beta, double* C, long long ldc,
           long long sC, long long batch, cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return SCIMLOP_OK;
  return scimlop_gemm_strided_batched(h, SCIMLOP_F64, SCIMLOP_PREC_DEFAULT, ta, tb, M, N, K, &alpha, A, lda, sA, B, ldb, sB,
                                      &beta, C, ldc, sC, batch, st);
}

// One triangular sweep along a mode: dst = T^{-1} src for the (Lp x n x R) view, T = L (forward) or U (backward).
// inner mode (Lp == 1): the arrays are n x R;  otherwise R slabs of Lp x n and the factor acts from the right.
int tri_sweep(scimlop_context* h, const Blob& f, long long n, bool upper, long long Lp, long long R, const double* src,
              double* dst, cudaStream_t st) {
  const long long nb = (n + DB - 1) / DB;
  const double* D = upper ? f.DU : f.DL;
  for (long long s = 0; s < nb; s++) {
    const long long i = upper ? nb - 1 - s : s;
    const long long o = i * DB, sz = std::min<long long>(DB, n - o);
    const long long ko = upper ? o + sz : 0, kw = upper ? n - ko : o;          // already-solved index range
    int rc;
    if (Lp == 1) {
      rc = gemm64(h, false, false, sz, R, sz, 1.0, D + o, f.ldf, 0, src + o, n, 0, 0.0, dst + o, n, 0, 1, st);
      if (!rc && kw > 0) rc = gemm64(h, false, false, sz, R, kw, -1.0, f.LUr + o + ko * f.ldf, f.ldf, 0, dst + ko, n, 0, 1.0, dst + o, n, 0, 1, st);
    } else {
      rc = gemm64(h, false, true, Lp, sz, sz, 1.0, src + Lp * o, Lp, Lp * n, D + o, f.ldf, 0, 0.0, dst + Lp * o, Lp, Lp * n, R, st);
      if (!rc && kw > 0)
        rc = gemm64(h, false, true, Lp, sz, kw, -1.0, dst + Lp * ko, Lp, Lp * n, f.LUr + o + ko * f.ldf, f.ldf, 0, 1.0, dst + Lp * o, Lp,
                    Lp * n, R, st);
    }
    if (rc) return rc;
  }
  return SCIMLOP_OK;
}

// The sweeps of the TRANSPOSED solve (A^T x = b: U^T y = b forward, L^T z = y backward).  dst = T^{-T} src with
// T = U (`unit_lower` false: forward over the lower-triangular U^T) or T = L (true: backward over the unit upper L^T).
// The pre-multiplied block rows of the blob belong to the other orientation, so a block step is the two GEMMs in the
// other order, on the plain factors:  t_i = b_i - T_{solved, i}^T y_solved  (INTO src: the sweep consumes its input),
// then  y_i = Dinv_ii^T t_i  into dst.  Same flops, same stability argument (inverted diagonal blocks only).
int tri_sweep_t(scimlop_context* h, const Blob& f, long long n, bool unit_lower, long long Lp, long long R, double* src,
                double* dst, cudaStream_t st) {
  const long long nb = (n + DB - 1) / DB;
  const double* D = unit_lower ? f.DL : f.DU;
  for (long long s = 0; s < nb; s++) {
    const long long i = unit_lower ? nb - 1 - s : s;
    const long long o = i * DB, sz = std::min<long long>(DB, n - o);
    const long long ko = unit_lower ? o + sz : 0, kw = unit_lower ? n - ko : o;   // already-solved index range
    const double* M = f.LU + ko + o * f.ldf;                                      // T_{solved, i}: kw x sz
    int rc = SCIMLOP_OK;
    if (Lp == 1) {
      if (kw > 0) rc = gemm64(h, true, false, sz, R, kw, -1.0, M, f.ldf, 0, dst + ko, n, 0, 1.0, src + o, n, 0, 1, st);
      if (!rc) rc = gemm64(h, true, false, sz, R, sz, 1.0, D + o, f.ldf, 0, src + o, n, 0, 0.0, dst + o, n, 0, 1, st);
    } else {
      if (kw > 0)
        rc = gemm64(h, false, false, Lp, sz, kw, -1.0, dst + Lp * ko, Lp, Lp * n, M, f.ldf, 0, 1.0, src + Lp * o, Lp, Lp * n, R, st);
      if (!rc) rc = gemm64(h, false, false, Lp, sz, sz, 1.0, src + Lp * o, Lp, Lp * n, D + o, f.ldf, 0, 0.0, dst + Lp * o, Lp, Lp * n, R, st);
    }
    if (rc) return rc;
  }
  return SCIMLOP_OK;
}

inline unsigned grid_for(long long total, scimlop_context* h) {
  return (unsigned)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16));
}

}  // namespace

extern "C" int scimlop_factorize_bytes(int dtype, int64_t n, size_t* fact_bytes, size_t* scratch) {
  if (dtype != SCIMLOP_F64) return SCIMLOP_ERR_UNSUPPORTED;
  if (n < 0) return SCIMLOP_ERR_BAD_SHAPE;
  if (fact_bytes) *fact_bytes = blob_bytes(n);
  if (scratch) *scratch = scratch_bytes(n);
  return SCIMLOP_OK;
}

extern "C" int scimlop_factorize(scimlop_handle_t h, int dtype, int64_t n, const void* A, int64_t lda, void* fact,
                                 size_t fact_bytes, void* Ainv, int64_t ldinv, void* scratch, size_t scratch_len,
                                 double* cond1, void* stream) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (dtype != SCIMLOP_F64) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "factorize: Float64 only (Float32 factors keep scimlop_invert)");
  if (n < 0 || lda < std::max<int64_t>(1, n) || (Ainv && ldinv < std::max<int64_t>(1, n)))
    return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "factorize: bad n / lda / ldinv");
  if (n > 32768) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "factorize: n = %lld above the 32768 rows one panel launch keeps co-resident", (long long)n);
  if (n == 0) { if (cond1) *cond1 = 0.0; return SCIMLOP_OK; }
  if (!A || !fact || !scratch) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "factorize: null pointer");
  if (fact_bytes < blob_bytes(n) || scratch_len < scratch_bytes(n))
    return scimlop_fail(h, SCIMLOP_ERR_WORKSPACE, "factorize: buffers too small (scimlop_factorize_bytes)");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Blob f = blob_at(fact, n);
  const Scratch s = scratch_at(scratch, n);
  const long long npan = (n + PB - 1) / PB;
  // clusters of up to 16 CTAs (non-portable size: opt in; fall back to the grid-barrier panel kernel if refused)
  static int cluster_fits[17] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
  const bool cluster16_ok = cudaFuncSetAttribute(lu_panel_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
  if (!cluster16_ok) cudaGetLastError();
  SCIMLOP_CUDA(h, cudaMemsetAsync(s.bar, 0, (size_t)npan * 4, st));
  SCIMLOP_CUDA(h, cudaMemsetAsync(s.norms, 0, 256, st));
  SCIMLOP_CUDA(h, cudaMemsetAsync(s.info, 0, 256, st));
  copy_matrix_kernel<<<grid_for(n * n, h), 256, 0, st>>>(n, static_cast<const double*>(A), lda, f.LU, f.ldf);
  norm1_kernel<<<(unsigned)std::min<long long>(n, 1024), 256, 0, st>>>(n, f.LU, f.ldf, s.norms);
  h->launches += 2;
#ifdef SCIMLOP_LU_TRACE
  // timing build (tools/lu_trace.py): timed events at the phase boundaries of every panel step
  static cudaEvent_t tev[1024][7];
  static bool tev_ready = false;
  if (!tev_ready) { for (auto& r : tev) for (auto& e : r) cudaEventCreate(&e); tev_ready = true; }
#define LU_TEV(p, i, stream) do { if ((p) < 1024) cudaEventRecord(tev[p][i], stream); } while (0)
#else
#define LU_TEV(p, i, stream) do { } while (0)
#endif
  // ---- blocked LU with look-ahead ----
  // Panel p + 1 (16 CTAs, ~100 us of barriers at n = 4096: the critical path) runs on the caller's stream right after the
  // updates of ITS columns; panel p's swaps, block-row solve and trailing update of every other column run beside it on
  // the handle's side stream (sequentially they were 11 of 30 ms at n = 4096).
  if (int rc = scimlop_ensure_side_stream(h)) return rc;
  cudaStream_t sB = h->s_side;
  SCIMLOP_CUDA(h, cudaEventRecord(h->ev_pipe[0], st));
  SCIMLOP_CUDA(h, cudaStreamWaitEvent(sB, h->ev_pipe[0], 0));
  int prev_j0 = 0, prev_w = 0;
  for (long long p = 0; p < npan; p++) {
    const int j0 = (int)(p * PB), w = (int)std::min<long long>(PB, n - j0);
    const unsigned ncta = (unsigned)((n - j0 + PROWS - 1) / PROWS);
    bool clustered = false;
    LU_TEV(p, 0, st);                              // panel start
    if (ncta <= 16 && (ncta <= 8 || cluster16_ok)) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(ncta, 1, 1);
      cfg.blockDim = dim3(PROWS, 1, 1);
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (cluster_fits[ncta] < 0) {       // once per cluster size: can the device place such a cluster at all?
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, lu_panel_cluster_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); nc = 0; }
        cluster_fits[ncta] = nc > 0 ? 1 : 0;
      }
      if (cluster_fits[ncta] == 1) {
        SCIMLOP_CUDA(h, cudaLaunchKernelEx(&cfg, lu_panel_cluster_kernel, (long long)n, f.LU, f.ldf, j0, w, s.piv, s.info, prev_j0, prev_w));
        clustered = true;
      }
    }
    if (!clustered) {
      lu_panel_kernel<<<ncta, PROWS, 0, st>>>(n, f.LU, f.ldf, j0, w, s.piv, s.bar + p, s.slots, s.info, prev_j0, prev_w);
    }
    h->launches++;
    // ---- look-ahead: the next panel's columns on this stream, every other column on the side stream ----
    const long long j1 = j0 + w, rem = n - j1;
    const int wn = (int)std::min<long long>(PB, rem);              // width of the next panel (0 after the last)
    if (n - w > 0) {
      cudaEvent_t evP = h->ev_pipe[0], evB = h->ev_pipe[1 + (p & 1)], evBprev = h->ev_pipe[1 + ((p + 1) & 1)];
      SCIMLOP_CUDA(h, cudaEventRecord(evP, st));
      LU_TEV(p, 1, st);                            // panel end
      if (wn > 0) {
        if (p > 0) SCIMLOP_CUDA(h, cudaStreamWaitEvent(st, evBprev, 0));   // those columns carry every earlier update
        lu_lookahead_prep_kernel<<<1, LA_THREADS, 0, st>>>(f.LU, f.ldf, j0, w, (int)j1, wn, s.piv);
        LU_TEV(p, 6, st);                          // swaps + block-row solve of the next panel's columns done
        h->launches++;                             // (the rank-w update of those columns rides in the next panel kernel)
      }
      LU_TEV(p, 2, st);                            // next panel's columns ready
      SCIMLOP_CUDA(h, cudaStreamWaitEvent(sB, evP, 0));
      LU_TEV(p, 3, sB);                            // side stream released
      const long long ncolsB = n - w - wn;
      if (ncolsB > 0) {
        lu_swap_trsm_kernel<<<(unsigned)((ncolsB + ST_THREADS - 1) / ST_THREADS), ST_THREADS, 0, sB>>>(n, f.LU, f.ldf, j0, w, s.piv, ncolsB, j0,
                                                                                                   (long long)w + wn);
        h->launches++;
      }
      LU_TEV(p, 4, sB);                            // swaps + block-row solve done
      const long long j2 = j1 + wn, remB = n - j2;
      if (rem > 0 && remB > 0) {
        // leave the SMs of the next panel's cluster free (the panel kernel is the critical path)
        const int nsm = std::max(1, h->sm_count - h->sm_reserve);
        const int next_ctas = (int)std::min<long long>((rem + PROWS - 1) / PROWS, nsm / 2);
        h->grid_limit = std::max(1, nsm - next_ctas);
        const int rc = gemm64(h, false, false, rem, remB, w, -1.0, f.LU + j1 + (long long)j0 * f.ldf, f.ldf, 0, f.LU + j0 + j2 * f.ldf, f.ldf, 0,
                              1.0, f.LU + j1 + j2 * f.ldf, f.ldf, 0, 1, sB);
        h->grid_limit = 0;
        if (rc) return rc;
      }
      SCIMLOP_CUDA(h, cudaEventRecord(evB, sB));
      LU_TEV(p, 5, sB);                            // trailing update done
    }
    prev_j0 = j0; prev_w = w;
  }
  if (npan > 0 && n > PB) SCIMLOP_CUDA(h, cudaStreamWaitEvent(st, h->ev_pipe[1 + ((npan - 1) & 1)], 0));   // join
  SCIMLOP_CUDA(h, cudaGetLastError());
  SCIMLOP_CUDA(h, cudaFuncSetAttribute(perm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
  // (a serial pass of one thread, ~0.2 ms at n = 4096: on the side stream, beside the inversion of the diagonal blocks)
  SCIMLOP_CUDA(h, cudaEventRecord(h->ev_pipe[0], st));
  SCIMLOP_CUDA(h, cudaStreamWaitEvent(sB, h->ev_pipe[0], 0));
  perm_kernel<<<1, 256, (size_t)n * 4, sB>>>((int)n, s.piv, f.perm);
  SCIMLOP_CUDA(h, cudaEventRecord(h->ev_pipe[1], sB));
  // ---- solve structures: inverse diagonal blocks, pre-multiplied off-diagonal block rows ----
  const int nb = (int)((n + DB - 1) / DB);
  SCIMLOP_CUDA(h, cudaFuncSetAttribute(tri_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DB * (DB + 1) * 8));
  tri_inv_kernel<<<2 * nb, DB, DB * (DB + 1) * 8, st>>>(n, f.LU, f.ldf, f.DL, f.DU, nb);
  h->launches += 2;
  for (int i = 0; i < nb; i++) {
    const long long o = (long long)i * DB, sz = std::min<long long>(DB, n - o);
    if (o > 0)
      if (int rc = gemm64(h, false, false, sz, o, sz, 1.0, f.DL + o, f.ldf, 0, f.LU + o, f.ldf, 0, 0.0, f.LUr + o, f.ldf, 0, 1, st)) return rc;
    const long long ko = o + sz, kw = n - ko;
    if (kw > 0)
      if (int rc = gemm64(h, false, false, sz, kw, sz, 1.0, f.DU + o, f.ldf, 0, f.LU + o + ko * f.ldf, f.ldf, 0, 0.0, f.LUr + o + ko * f.ldf,
                          f.ldf, 0, 1, st))
        return rc;
  }
  // ---- explicit inverse = the solve applied to the identity; cond_1 ----
  SCIMLOP_CUDA(h, cudaStreamWaitEvent(st, h->ev_pipe[1], 0));                       // perm is ready
  identity_perm_kernel<<<grid_for(n * n, h), 256, 0, st>>>(n, n, f.perm, s.T1);     // T1 = P (n x n, ld n)
  h->launches++;
  if (int rc = tri_sweep(h, f, n, false, 1, n, s.T1, s.T2, st)) return rc;
  if (int rc = tri_sweep(h, f, n, true, 1, n, s.T2, s.T1, st)) return rc;
  if (Ainv) {
    copy_matrix_kernel<<<grid_for(n * n, h), 256, 0, st>>>(n, s.T1, n, static_cast<double*>(Ainv), ldinv);
    h->launches++;
  }
  norm1_kernel<<<(unsigned)std::min<long long>(n, 1024), 256, 0, st>>>(n, s.T1, n, s.norms + 1);
  h->launches++;
  SCIMLOP_CUDA(h, cudaGetLastError());
  double hn[2];
  int hinfo = 0;
  SCIMLOP_CUDA(h, cudaMemcpyAsync(hn, s.norms, 16, cudaMemcpyDeviceToHost, st));
  SCIMLOP_CUDA(h, cudaMemcpyAsync(&hinfo, s.info, 4, cudaMemcpyDeviceToHost, st));
  SCIMLOP_CUDA(h, cudaStreamSynchronize(st));
#ifdef SCIMLOP_LU_TRACE
  for (long long p = 0; p + 1 < npan && p < 1024; p += (npan > 16 ? npan / 16 : 1)) {
    float t[7];
    for (int i = 0; i < 7; i++) cudaEventElapsedTime(&t[i], tev[0][0], tev[p][i]);
    fprintf(stderr, "lu_trace n=%lld panel %3lld: P %8.1f..%8.1f us | prep %8.1f cols ready %8.1f | side: released %8.1f swaps %8.1f update %8.1f\n",
            (long long)n, p, t[0] * 1e3, t[1] * 1e3, t[6] * 1e3, t[2] * 1e3, t[3] * 1e3, t[4] * 1e3, t[5] * 1e3);
  }
  { float t; cudaEventElapsedTime(&t, tev[0][0], tev[std::min<long long>(npan, 1024) - 2][5]); fprintf(stderr, "lu_trace last side-stream update ends at %.1f us\n", t * 1e3); }
#endif
  if (hinfo != 0) return scimlop_fail(h, SCIMLOP_ERR_SINGULAR, "factorize: zero pivot at column %d", hinfo);
  if (cond1) *cond1 = hn[0] * hn[1];
  return SCIMLOP_OK;
}

// W = (F_1 (x) ... (x) F_n)^{-1} V with every dense factor given by its factorisation blob (kinds: SCIMLOP_DENSE,
// SCIMLOP_DENSE | SCIMLOP_TRANS for the transposed factor, or SCIMLOP_IDENTITY; dims[f] = order of factor f, outermost first).  workspace: one buffer of prod(dims) * k doubles.
extern "C" int scimlop_kron_ldiv(scimlop_handle_t h, int dtype, int nfactors, void* const* facts, const int64_t* dims,
                                 const int32_t* kinds, const void* V, void* W, int64_t k, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (dtype != SCIMLOP_F64) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "kron_ldiv: Float64 only");
  if (nfactors < 1 || nfactors > 8 || !dims || !kinds || !facts) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_ldiv: bad factor description");
  long long N = 1;
  int nops = 0;
  for (int f = 0; f < nfactors; f++) {
    if (dims[f] < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kron_ldiv: negative size");
    if ((kinds[f] & ~SCIMLOP_TRANS) != SCIMLOP_DENSE && kinds[f] != SCIMLOP_IDENTITY)
      return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "kron_ldiv: factor %d: only dense (factorised) and identity factors", f);
    if (kinds[f] != SCIMLOP_IDENTITY) { nops++; if (!facts[f]) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_ldiv: factor %d has no factorisation", f); }
    N *= dims[f];
  }
  if (k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kron_ldiv: negative k");
  if (k == 0 || N == 0) return SCIMLOP_OK;
  if (!V || !W) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_ldiv: V/W null");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bytes = (size_t)N * (size_t)k * 8;
  if (nops == 0) {
    SCIMLOP_CUDA(h, cudaMemcpyAsync(W, V, bytes, cudaMemcpyDeviceToDevice, st));
    return SCIMLOP_OK;
  }
  if (!workspace || workspace_bytes < bytes) return scimlop_fail(h, SCIMLOP_ERR_WORKSPACE, "kron_ldiv: workspace %zu B < required %zu B", workspace_bytes, bytes);
  // three out-of-place passes per factor -- row gather, L sweep, U sweep; for a transposed factor (SCIMLOP_TRANS: the
  // lazy adjoint, A^T = U^T L^T P) U^T sweep, L^T sweep, row scatter -- alternate between W and the workspace and must
  // end in W.  The transposed sweeps consume their source, so if the first factor applied is transposed V is copied first.
  int first_op = -1;
  for (int fi = nfactors - 1; fi >= 0 && first_op < 0; fi--)
    if (kinds[fi] != SCIMLOP_IDENTITY) first_op = fi;
  const bool copy_first = (kinds[first_op] & SCIMLOP_TRANS) != 0;
  const int passes = 3 * nops + (copy_first ? 1 : 0);
  double* bufs[2] = {static_cast<double*>(W), static_cast<double*>(workspace)};
  int cur = passes % 2 == 1 ? 0 : 1;              // target of the first pass
  const double* src = static_cast<const double*>(V);
  if (copy_first) {
    SCIMLOP_CUDA(h, cudaMemcpyAsync(bufs[cur], V, bytes, cudaMemcpyDeviceToDevice, st));
    src = bufs[cur];
    cur ^= 1;
  }
  long long Lp = 1;
  for (int fi = nfactors - 1; fi >= 0; fi--) {
    const long long n = dims[fi];
    if (kinds[fi] == SCIMLOP_IDENTITY) { Lp *= n; continue; }
    const long long R = N * k / (Lp * n);
    const Blob f = blob_at(facts[fi], n);
    if (kinds[fi] & SCIMLOP_TRANS) {
      // src is bufs[cur ^ 1] here (a buffer, never V): sweep into bufs[cur], back into bufs[cur ^ 1], scatter into bufs[cur]
      double* a = bufs[cur ^ 1];
      if (int rc = tri_sweep_t(h, f, n, false, Lp, R, a, bufs[cur], st)) return rc;
      if (int rc = tri_sweep_t(h, f, n, true, Lp, R, bufs[cur], a, st)) return rc;
      perm_mode_scatter_kernel<<<grid_for(N * k, h), 256, 0, st>>>(Lp, n, R, f.perm, a, bufs[cur]);
      h->launches++;
    } else {
      perm_mode_kernel<<<grid_for(N * k, h), 256, 0, st>>>(Lp, n, R, f.perm, src, bufs[cur]);
      h->launches++;
      if (int rc = tri_sweep(h, f, n, false, Lp, R, bufs[cur], bufs[cur ^ 1], st)) return rc;
      if (int rc = tri_sweep(h, f, n, true, Lp, R, bufs[cur ^ 1], bufs[cur], st)) return rc;
    }
    src = bufs[cur];
    cur ^= 1;
    Lp *= n;
  }
  SCIMLOP_CUDA(h, cudaGetLastError());
  return SCIMLOP_OK;
}
