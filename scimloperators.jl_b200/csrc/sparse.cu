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
        for (int c = 0; c < C; c++) acc[c] += a * __ldg(vp + c * ldv);
      }
    } else {
      for (int j = j0; j < j1; j++) {
        const T a = vals[j];
        const T* vp = v0 + (colidx[j] - base);
#pragma unroll
        for (int c = 0; c < C; c++)
          if (c0 + c < k) acc[c] += a * __ldg(vp + c * ldv);
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
        if (c0 + c < k) atomicAdd(wp + c * ldw, a * x[c]);
    }
  }
}

template <typename T>
int csr_mul_t(scimlop_context* h, bool trans, long long m, long long n, long long k, const int* rowptr, const int* colidx,
              const T* vals, int base, const T* V, long long ldv, T* W, long long ldw, T alpha, T beta, cudaStream_t st) {
  constexpr int THREADS = 128;
  const unsigned gx = (unsigned)((m + THREADS - 1) / THREADS);
  if (!trans) {
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

}  // namespace

extern "C" int scimlop_csr_mul(scimlop_handle_t h, int dtype, int trans, int64_t rows, int64_t cols, int64_t k,
                               const int32_t* rowptr, const int32_t* colidx, const void* values, int index_base,
                               const void* V, int64_t ldv, void* W, int64_t ldw, const void* alpha, const void* beta,
                               void* stream) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "csr_mul: unsupported dtype %d", dtype);
  if (rows < 0 || cols < 0 || k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "csr_mul: negative size");
  if (index_base != 0 && index_base != 1) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "csr_mul: index_base must be 0 or 1");
  const int64_t out_rows = trans ? cols : rows, in_rows = trans ? rows : cols;
  if (ldv < std::max<int64_t>(1, in_rows) || ldw < std::max<int64_t>(1, out_rows))
    return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "csr_mul: ld smaller than the rows");
  if (k == 0 || out_rows == 0) return SCIMLOP_OK;
  if (!W || (rows > 0 && !rowptr)) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "csr_mul: null pointer");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SCIMLOP_F64) {
    const double a = alpha ? *static_cast<const double*>(alpha) : 1.0, b = beta ? *static_cast<const double*>(beta) : 0.0;
    return csr_mul_t<double>(h, trans != 0, rows, cols, k, rowptr, colidx, static_cast<const double*>(values), index_base,
                             static_cast<const double*>(V), ldv, static_cast<double*>(W), ldw, a, b, st);
  }
  const float a = alpha ? *static_cast<const float*>(alpha) : 1.f, b = beta ? *static_cast<const float*>(beta) : 0.f;
  return csr_mul_t<float>(h, trans != 0, rows, cols, k, rowptr, colidx, static_cast<const float*>(values), index_base,
                          static_cast<const float*>(V), ldv, static_cast<float*>(W), ldw, a, b, st);
}
