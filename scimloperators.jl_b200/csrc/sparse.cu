// sparse.cu -- MatrixOperator over a general sparse matrix (CSR): W(m x k) = alpha * op(A) * V + beta * W.
//
// The reference reaches this through `mul!(w, L.A, v[, α, β])` (src/matrix.jl:337-350) with L.A a SparseMatrixCSC
// (test/basic.jl:405-426 and test/Downstream/alloccheck.jl:37-45 build their operators with `sprand`;
// ext/SciMLOperatorsSparseArraysExt.jl:8-23 converts operators to sparse): the arithmetic is SparseArrays' stdlib
// SpMM, a gather per output row for A in row-compressed form and a scatter per input row for the transposed form.
// On the device the matrix lives as CSR (CUDA.jl: CuSparseMatrixCSR, 1-based Int32 indices -> `index_base`); a
// SparseMatrixCSC is the CSR form of its transpose, which is what `trans` is for.
//
// Gather kernel: one thread per (row, tile of C columns).  Consecutive threads own consecutive rows, so W is written
// with full 128-byte lines per column and each thread walks its own slice of (colidx, values) -- slices of neighbouring
// rows are adjacent in memory, every line of them is used completely (through L1).  The C accumulators amortise the
// index / value loads; V is gathered per column (a column of V is n * 8 bytes: L2-resident for the operator sizes
// this path is for).  HBM-bound on (1 + k / C-reuse) * nnz * 12 B + (n + m) * k * 8 B.
// Scatter kernel (trans): W is pre-scaled by beta, then every nonzero adds alpha * a * V(i, c) into W(col, c) with
// atomicAdd -- the summation order, hence the last bits, vary from run to run (stated in the header).
#include "context.h"

namespace {

// ComplexF32 / ComplexF64 values (interleaved re, im): the `sprand(ComplexF64, ...)` operators of alloccheck.jl:37-45
template <typename R>
struct cplx {
  R re, im;
  __host__ __device__ constexpr cplx(R r = R(0), R i = R(0)) : re(r), im(i) {}
  __host__ __device__ cplx operator*(const cplx& o) const { return cplx(re * o.re - im * o.im, re * o.im + im * o.re); }
  __host__ __device__ cplx operator+(const cplx& o) const { return cplx(re + o.re, im + o.im); }
  __host__ __device__ cplx& operator+=(const cplx& o) { re += o.re; im += o.im; return *this; }
  __host__ __device__ bool operator!=(const cplx& o) const { return re != o.re || im != o.im; }
  __host__ __device__ bool operator==(const cplx& o) const { return re == o.re && im == o.im; }
};
template <typename T> __device__ __forceinline__ T ldg_any(const T* p) { return __ldg(p); }
template <> __device__ __forceinline__ cplx<double> ldg_any(const cplx<double>* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return cplx<double>(v.x, v.y);
}
template <> __device__ __forceinline__ cplx<float> ldg_any(const cplx<float>* p) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return cplx<float>(v.x, v.y);
}
template <typename T> __device__ __forceinline__ void atomic_add_any(T* p, T v) { atomicAdd(p, v); }
template <typename R> __device__ __forceinline__ void atomic_add_any(cplx<R>* p, cplx<R> v) {
  atomicAdd(&p->re, v.re);
  atomicAdd(&p->im, v.im);
}
template <typename T> struct is_cplx { static constexpr bool value = false; };
template <typename R> struct is_cplx<cplx<R>> { static constexpr bool value = true; };

template <typename T, int C>
__global__ void csr_gather_kernel(long long m, long long k, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                  const T* __restrict__ vals, int base, const T* __restrict__ V, long long ldv, T* W,
                                  long long ldw, T alpha, T beta) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m) return;
  const int j0 = rowptr[row] - base, j1 = rowptr[row + 1] - base;
  for (long long c0 = (long long)blockIdx.y * C; c0 < k; c0 += (long long)gridDim.y * C) {
    T acc[C];
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = T(0);
    const T* v0 = V + c0 * ldv;
    if (c0 + C <= k) {
      for (int j = j0; j < j1; j++) {
        const T a = vals[j];
        const T* vp = v0 + (colidx[j] - base);
#pragma unroll
        for (int c = 0; c < C; c++) acc[c] += a * ldg_any(vp + c * ldv);
      }
    } else {
      for (int j = j0; j < j1; j++) {
        const T a = vals[j];
        const T* vp = v0 + (colidx[j] - base);
#pragma unroll
        for (int c = 0; c < C; c++)
          if (c0 + c < k) acc[c] += a * ldg_any(vp + c * ldv);
      }
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
      if (c0 + c < k) {
        T* w = W + row + (c0 + c) * ldw;
        T r = alpha * acc[c];
        if (beta != T(0)) r += beta * *w;     // beta == 0: W is never read (NaN-prefilled outputs stay clean)
        *w = r;
      }
    }
  }
}

// V (n x k, column-major, ldv) -> Vt (k x n: the k entries of a row contiguous), 32 x 32 tiles through shared memory
template <typename T>
__global__ void transpose_kernel(long long n, long long k, const T* __restrict__ V, long long ldv, T* __restrict__ Vt) {
  __shared__ T tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32, c0 = (long long)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long r = r0 + threadIdx.x, c = c0 + j;
    tile[j][threadIdx.x] = (r < n && c < k) ? V[r + c * ldv] : T(0);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long r = r0 + j, c = c0 + threadIdx.x;
    if (r < n && c < k) Vt[c + r * k] = tile[threadIdx.x][j];
  }
}

// Gather from the row-contiguous copy: a nonzero costs C contiguous values (full 32-byte sectors) instead of C sectors
// of which one word each is used -- 4x less L2 sector traffic for Float64, 8x for Float32.
template <typename T, int C>
__global__ void csr_gather_t_kernel(long long m, long long k, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                    const T* __restrict__ vals, int base, const T* __restrict__ Vt, T* W, long long ldw,
                                    T alpha, T beta) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m) return;
  const int j0 = rowptr[row] - base, j1 = rowptr[row + 1] - base;
  for (long long c0 = (long long)blockIdx.y * C; c0 < k; c0 += (long long)gridDim.y * C) {   // k % C == 0 on this path
    T acc[C];
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = T(0);
    for (int j = j0; j < j1; j++) {
      const T a = vals[j];
      const T* vp = Vt + (long long)(colidx[j] - base) * k + c0;
      T x[C];
#pragma unroll
      for (int c = 0; c < C; c += 16 / (int)sizeof(T)) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(vp + c));
        *reinterpret_cast<float4*>(&x[c]) = q;
      }
#pragma unroll
      for (int c = 0; c < C; c++) acc[c] += a * x[c];
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
      T* w = W + row + (c0 + c) * ldw;
      T r = alpha * acc[c];
      if (beta != T(0)) r += beta * *w;
      *w = r;
    }
  }
}

// The same gather with one WARP per row and the lanes across columns: a nonzero costs the warp ONE fully coalesced read
// of 512 contiguous bytes of the row-contiguous copy (the thread-per-row form above touches 32 different lines per load
// instruction, half a sector used each: 3.5 TB/s of L2 traffic, 89 % of the stalls on the scoreboard).  The (index,
// value) pairs of a row are loaded 32 at a time, one per lane, and broadcast by shuffle, so the gathers of a row do not
// depend on each other: 8 in flight per warp.  A CTA owns 32 rows x 32 * VEC columns; the tile goes through shared
// memory so that W (column-major) is written with 256 contiguous bytes per column.
template <typename T, int NV>
__global__ void __launch_bounds__(256) csr_gather_rows_kernel(long long m, long long k, const int* __restrict__ rowptr,
                                                              const int* __restrict__ colidx, const T* __restrict__ vals, int base,
                                                              const T* __restrict__ Vt, T* W, long long ldw, T alpha, T beta) {
  // NV 16-byte loads per lane and nonzero (NV = 2: a 1 KB run per nonzero, half the index / shuffle / address work per
  // byte: ncu showed the NV = 1 form 56 % issue-bound with the broadcast shuffle as its hottest line)
  constexpr int VEC = 16 / (int)sizeof(T);       // columns per 16-byte load
  constexpr int CB = 32 * VEC * NV;              // columns per CTA
  __shared__ T tile[32][CB + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * 32;
  for (long long c0 = (long long)blockIdx.y * CB; c0 < k; c0 += (long long)gridDim.y * CB) {   // k % CB == 0 on this path
    for (int rr = 0; rr < 4; rr++) {
      const long long r = row0 + warp + 8 * rr;
      T acc[NV][VEC];
#pragma unroll
      for (int v = 0; v < NV; v++)
#pragma unroll
        for (int e = 0; e < VEC; e++) acc[v][e] = T(0);
      if (r < m) {
        const int j0 = rowptr[r] - base, j1 = rowptr[r + 1] - base;
        const T* vbase = Vt + c0 + lane * VEC;
        for (int jb = j0; jb < j1; jb += 32) {
          const int cnt = min(32, j1 - jb);
          const int myc = lane < cnt ? colidx[jb + lane] - base : 0;
          const T mya = lane < cnt ? vals[jb + lane] : T(0);
          constexpr int UQ = 8 / NV;             // nonzeros in flight: 8 loads per lane either way
          for (int q = 0; q < cnt; q += UQ) {
            float4 x[UQ][NV];
            T a[UQ];
#pragma unroll
            for (int u = 0; u < UQ; u++) {
              const int cu = __shfl_sync(0xffffffffu, myc, (q + u) & 31);
              a[u] = __shfl_sync(0xffffffffu, mya, (q + u) & 31);
              const T* xp = vbase + (long long)cu * k;
#pragma unroll
              for (int v = 0; v < NV; v++) {
                if (q + u < cnt) x[u][v] = __ldg(reinterpret_cast<const float4*>(xp + v * 32 * VEC));
                else x[u][v] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              if (q + u >= cnt) a[u] = T(0);
            }
#pragma unroll
            for (int u = 0; u < UQ; u++)
#pragma unroll
              for (int v = 0; v < NV; v++) {
                const T* xv = reinterpret_cast<const T*>(&x[u][v]);
#pragma unroll
                for (int e = 0; e < VEC; e++) acc[v][e] += a[u] * xv[e];
              }
          }
        }
      }
#pragma unroll
      for (int v = 0; v < NV; v++)
#pragma unroll
        for (int e = 0; e < VEC; e++) tile[warp + 8 * rr][(v * 32 + lane) * VEC + e] = acc[v][e];
    }
    __syncthreads();
    const long long r = row0 + lane;
    if (r < m) {
      for (int c = warp; c < CB; c += 8) {
        T* w = W + r + (c0 + c) * ldw;
        T v = alpha * tile[lane][c];
        if (beta != T(0)) v += beta * *w;
        *w = v;
      }
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void scale_cols_kernel(long long n, long long k, T* W, long long ldw, T beta) {
  const long long total = n * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    T* w = W + (i % n) + (i / n) * ldw;
    *w = (beta == T(0)) ? T(0) : beta * *w;
  }
}

template <typename T, int C>
__global__ void csr_scatter_kernel(long long m, long long k, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                   const T* __restrict__ vals, int base, const T* __restrict__ V, long long ldv, T* W,
                                   long long ldw, T alpha) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // row of the STORED matrix = input index
  if (row >= m) return;
  const int j0 = rowptr[row] - base, j1 = rowptr[row + 1] - base;
  for (long long c0 = (long long)blockIdx.y * C; c0 < k; c0 += (long long)gridDim.y * C) {
    T x[C];
#pragma unroll
    for (int c = 0; c < C; c++) x[c] = (c0 + c < k) ? alpha * V[row + (c0 + c) * ldv] : T(0);
    for (int j = j0; j < j1; j++) {
      const T a = vals[j];
      T* wp = W + (colidx[j] - base) + c0 * ldw;
#pragma unroll
      for (int c = 0; c < C; c++)
        if (c0 + c < k) atomic_add_any(wp + c * ldw, a * x[c]);
    }
  }
}

template <typename T>
int csr_mul_t(scimlop_context* h, bool trans, long long m, long long n, long long k, const int* rowptr, const int* colidx,
              const T* vals, int base, const T* V, long long ldv, T* W, long long ldw, T alpha, T beta, void* workspace,
              size_t ws_bytes, cudaStream_t st) {
  constexpr int THREADS = 128;
  const unsigned gx = (unsigned)((m + THREADS - 1) / THREADS);
  if (!trans) {
    if (!is_cplx<T>::value && workspace && k >= 8 && k % 8 == 0 && ws_bytes >= (size_t)n * (size_t)k * sizeof(T) &&
        (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0 && (n + 31) / 32 <= 2147483647LL && (k + 31) / 32 <= 65535) {
      // enough columns to pay for a transposed staging copy of V (one extra read + write of V)
      T* Vt = static_cast<T*>(workspace);
      transpose_kernel<T><<<dim3((unsigned)((n + 31) / 32), (unsigned)((k + 31) / 32)), dim3(32, 8), 0, st>>>(n, k, V, ldv, Vt);
      bool by_rows = false;
      if constexpr (!is_cplx<T>::value) {
        const long long run = 512 / (long long)sizeof(T);   // columns of one 512-byte coalesced warp load
        const unsigned gr = (unsigned)((m + 31) / 32);
        if (k % (2 * run) == 0) {          // a warp per row, two 512-byte gathers per nonzero
          csr_gather_rows_kernel<T, 2><<<dim3(gr, (unsigned)std::min<long long>(k / (2 * run), 65535)), 256, 0, st>>>(
              m, k, rowptr, colidx, vals, base, Vt, W, ldw, alpha, beta);
          by_rows = true;
        } else if (k % run == 0) {
          csr_gather_rows_kernel<T, 1><<<dim3(gr, (unsigned)std::min<long long>(k / run, 65535)), 256, 0, st>>>(
              m, k, rowptr, colidx, vals, base, Vt, W, ldw, alpha, beta);
          by_rows = true;
        }
      }
      if (by_rows) {
      } else if (k % 16 == 0) {   // 16 columns per thread: a gathered row is a full 128-byte line (Float64)
        const unsigned gy = (unsigned)std::min<long long>(k / 16, 65535);
        csr_gather_t_kernel<T, 16><<<dim3(gx, gy), THREADS, 0, st>>>(m, k, rowptr, colidx, vals, base, Vt, W, ldw, alpha, beta);
      } else {
        const unsigned gy = (unsigned)std::min<long long>(k / 8, 65535);
        csr_gather_t_kernel<T, 8><<<dim3(gx, gy), THREADS, 0, st>>>(m, k, rowptr, colidx, vals, base, Vt, W, ldw, alpha, beta);
      }
      h->launches += 2;
      SCIMLOP_CUDA(h, cudaGetLastError());
      return SCIMLOP_OK;
    }
    if (k >= 8) {
      const unsigned gy = (unsigned)std::min<long long>((k + 7) / 8, 65535);
      csr_gather_kernel<T, 8><<<dim3(gx, gy), THREADS, 0, st>>>(m, k, rowptr, colidx, vals, base, V, ldv, W, ldw, alpha, beta);
    } else if (k >= 2) {
      csr_gather_kernel<T, 2><<<dim3(gx, (unsigned)((k + 1) / 2)), THREADS, 0, st>>>(m, k, rowptr, colidx, vals, base, V, ldv, W, ldw, alpha, beta);
    } else {
      csr_gather_kernel<T, 1><<<dim3(gx, 1), THREADS, 0, st>>>(m, k, rowptr, colidx, vals, base, V, ldv, W, ldw, alpha, beta);
    }
    h->launches++;
  } else {
    const long long total = n * k;
    const unsigned gs = (unsigned)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16));
    scale_cols_kernel<T><<<gs, 256, 0, st>>>(n, k, W, ldw, beta);
    const unsigned gy = (unsigned)std::min<long long>((k + 3) / 4, 65535);
    csr_scatter_kernel<T, 4><<<dim3(gx, gy), THREADS, 0, st>>>(m, k, rowptr, colidx, vals, base, V, ldv, W, ldw, alpha);
    h->launches += 2;
  }
  SCIMLOP_CUDA(h, cudaGetLastError());
  return SCIMLOP_OK;
}

// Sparse factor along an OUTER Kronecker mode: Y(l, i, r) = alpha * sum_j A(i, j) X(l, j, r) + beta * Y(l, i, r) for the
// (L x n x R) -> (L x m x R) views.  The L entries of one (j, r) are contiguous, which is exactly the row-contiguous form
// the staged gather wants: no copy.  One warp per (i, r), lanes across l (coalesced runs of 32 elements, 4 per lane in
// flight), the (index, value) pairs of row i loaded 32 at a time and broadcast by shuffle.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) csr_mode_kernel(long long L, long long m, long long n, long long R, const int* __restrict__ rowptr,
                                                       const int* __restrict__ colidx, const T* __restrict__ vals, int base,
                                                       const T* __restrict__ X, T* Y, T alpha, T beta) {
  // VEC elements per load: 16 bytes when L and the pointers allow (half the instructions per byte), else 1
  constexpr int U = 4 / (VEC > 1 ? 2 : 1);        // loads in flight per lane and nonzero: 4 scalar or 2 vector
  constexpr int SPAN = 32 * VEC * U;              // elements of a mode row one pass of the warp covers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * 8 + warp;
  if (i >= m) return;
  const int j0 = rowptr[i] - base, j1 = rowptr[i + 1] - base;
  for (long long r = blockIdx.y; r < R; r += gridDim.y) {
    const T* Xr = X + L * n * r;
    T* Yr = Y + L * (i + m * r);
    for (long long l0 = 0; l0 < L; l0 += SPAN) {
      T acc[U][VEC];
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int e = 0; e < VEC; e++) acc[u][e] = T(0);
      for (int jb = j0; jb < j1; jb += 32) {
        const int cnt = min(32, j1 - jb);
        const int myc = lane < cnt ? colidx[jb + lane] - base : 0;
        const T mya = lane < cnt ? vals[jb + lane] : T(0);
        for (int q = 0; q < cnt; q++) {
          const long long cq = __shfl_sync(0xffffffffu, myc, q);
          const T aq = __shfl_sync(0xffffffffu, mya, q);
          const T* xp = Xr + L * cq + l0 + lane * VEC;
#pragma unroll
          for (int u = 0; u < U; u++) {
            if (l0 + (lane + 32 * u) * VEC < L) {
              if constexpr (VEC > 1) {
                const float4 x4 = __ldg(reinterpret_cast<const float4*>(xp + 32 * VEC * u));
                const T* xv = reinterpret_cast<const T*>(&x4);
#pragma unroll
                for (int e = 0; e < VEC; e++) acc[u][e] += aq * xv[e];
              } else {
                acc[u][0] += aq * __ldg(xp + 32 * u);
              }
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const long long l = l0 + (lane + 32 * u) * VEC;
        if (l < L) {
#pragma unroll
          for (int e = 0; e < VEC; e++) {
            T v = alpha * acc[u][e];
            if (beta != T(0)) v += beta * Yr[l + e];
            Yr[l + e] = v;
          }
        }
      }
    }
  }
}

template <typename T>
int csr_mode_t(scimlop_context* h, const scimlop_csr_factor* d, long long L, long long m, long long n, long long R, const T* X, T* Y,
               T alpha, T beta, void* stage, size_t stage_bytes, cudaStream_t st) {
  if (!d || !d->rowptr || (!d->colidx && m > 0) ) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron: sparse factor descriptor / arrays null");
  if (d->index_base != 0 && d->index_base != 1) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kron: sparse factor index_base must be 0 or 1");
  if (m == 0 || R == 0 || L == 0) return SCIMLOP_OK;
  if (L == 1)   // innermost mode: the plain SpMM  Y(m x R) = A X(n x R)
    return csr_mul_t<T>(h, false, m, n, R, d->rowptr, d->colidx, static_cast<const T*>(d->values), d->index_base, X, n, Y, m, alpha, beta,
                        stage, stage_bytes, st);
  const unsigned gx = (unsigned)((m + 7) / 8), gy = (unsigned)std::min<long long>(R, 65535);
  constexpr int VMAX = 16 / (int)sizeof(T);
  if (L % VMAX == 0 && (reinterpret_cast<uintptr_t>(X) & 15u) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15u) == 0)
    csr_mode_kernel<T, VMAX><<<dim3(gx, gy), 256, 0, st>>>(L, m, n, R, d->rowptr, d->colidx, static_cast<const T*>(d->values), d->index_base,
                                                           X, Y, alpha, beta);
  else
    csr_mode_kernel<T, 1><<<dim3(gx, gy), 256, 0, st>>>(L, m, n, R, d->rowptr, d->colidx, static_cast<const T*>(d->values), d->index_base, X,
                                                        Y, alpha, beta);
  h->launches++;
  SCIMLOP_CUDA(h, cudaGetLastError());
  return SCIMLOP_OK;
}

}  // namespace

int scimlop_csr_mode_impl(scimlop_context* c, int dtype, const void* desc, long long L, long long m, long long n, long long R,
                          const void* X, void* Y, const void* alpha, const void* beta, void* stage, size_t stage_bytes,
                          cudaStream_t stream) {
  const scimlop_csr_factor* d = static_cast<const scimlop_csr_factor*>(desc);
  if (dtype == SCIMLOP_F64)
    return csr_mode_t<double>(c, d, L, m, n, R, static_cast<const double*>(X), static_cast<double*>(Y), *static_cast<const double*>(alpha),
                              *static_cast<const double*>(beta), stage, stage_bytes, stream);
  return csr_mode_t<float>(c, d, L, m, n, R, static_cast<const float*>(X), static_cast<float*>(Y), *static_cast<const float*>(alpha),
                           *static_cast<const float*>(beta), stage, stage_bytes, stream);
}

extern "C" int scimlop_csr_workspace_bytes(int dtype, int trans, int64_t rows, int64_t cols, int64_t k, size_t* bytes) {
  if (!bytes) return SCIMLOP_ERR_NULL_POINTER;
  if (dtype == SCIMLOP_C64 || dtype == SCIMLOP_C128) { *bytes = 0; return SCIMLOP_OK; }   // complex: the direct gather
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32) return SCIMLOP_ERR_UNSUPPORTED;
  if (rows < 0 || cols < 0 || k < 0) return SCIMLOP_ERR_BAD_SHAPE;
  // the transposed staging copy of V pays off from 8 columns on (non-transposed form, k a multiple of 8)
  *bytes = (!trans && k >= 8 && k % 8 == 0) ? (size_t)cols * (size_t)k * (dtype == SCIMLOP_F64 ? 8 : 4) : 0;
  return SCIMLOP_OK;
}

extern "C" int scimlop_csr_mul(scimlop_handle_t h, int dtype, int trans, int64_t rows, int64_t cols, int64_t k,
                               const int32_t* rowptr, const int32_t* colidx, const void* values, int index_base,
                               const void* V, int64_t ldv, void* W, int64_t ldw, const void* alpha, const void* beta,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32 && dtype != SCIMLOP_C64 && dtype != SCIMLOP_C128)
    return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "csr_mul: unsupported dtype %d", dtype);
  if (rows < 0 || cols < 0 || k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "csr_mul: negative size");
  if (index_base != 0 && index_base != 1) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "csr_mul: index_base must be 0 or 1");
  const int64_t out_rows = trans ? cols : rows, in_rows = trans ? rows : cols;
  if (ldv < std::max<int64_t>(1, in_rows) || ldw < std::max<int64_t>(1, out_rows))
    return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "csr_mul: ld smaller than the rows");
  if (k == 0 || out_rows == 0) return SCIMLOP_OK;
  if (!W || (rows > 0 && !rowptr)) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "csr_mul: null pointer");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SCIMLOP_C128 || dtype == SCIMLOP_C64) {       // (re, im) pairs; alpha / beta point to one pair each
    if (dtype == SCIMLOP_C128) {
      using Z = cplx<double>;
      const double* ap = static_cast<const double*>(alpha);
      const double* bp = static_cast<const double*>(beta);
      return csr_mul_t<Z>(h, trans != 0, rows, cols, k, rowptr, colidx, static_cast<const Z*>(values), index_base,
                          static_cast<const Z*>(V), ldv, static_cast<Z*>(W), ldw, ap ? Z(ap[0], ap[1]) : Z(1.0), bp ? Z(bp[0], bp[1]) : Z(0.0),
                          nullptr, 0, st);
    }
    using Z = cplx<float>;
    const float* ap = static_cast<const float*>(alpha);
    const float* bp = static_cast<const float*>(beta);
    return csr_mul_t<Z>(h, trans != 0, rows, cols, k, rowptr, colidx, static_cast<const Z*>(values), index_base,
                        static_cast<const Z*>(V), ldv, static_cast<Z*>(W), ldw, ap ? Z(ap[0], ap[1]) : Z(1.f), bp ? Z(bp[0], bp[1]) : Z(0.f),
                        nullptr, 0, st);
  }
  if (dtype == SCIMLOP_F64) {
    const double a = alpha ? *static_cast<const double*>(alpha) : 1.0, b = beta ? *static_cast<const double*>(beta) : 0.0;
    return csr_mul_t<double>(h, trans != 0, rows, cols, k, rowptr, colidx, static_cast<const double*>(values), index_base,
                             static_cast<const double*>(V), ldv, static_cast<double*>(W), ldw, a, b, workspace, workspace_bytes, st);
  }
  const float a = alpha ? *static_cast<const float*>(alpha) : 1.f, b = beta ? *static_cast<const float*>(beta) : 0.f;
  return csr_mul_t<float>(h, trans != 0, rows, cols, k, rowptr, colidx, static_cast<const float*>(values), index_base,
                          static_cast<const float*>(V), ldv, static_cast<float*>(W), ldw, a, b, workspace, workspace_bytes, st);
}
