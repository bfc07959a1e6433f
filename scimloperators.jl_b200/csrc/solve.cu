// solve.cu -- the ldiv! side of the path.
//
// The reference solves with a cached factorisation: `factorize(L::TensorProductOperator)` factorises every
// factor (src/tensor.jl:227), `ldiv!(w, L, v)` then runs `ldiv!(C5, inner, U)` and `outer_div!`
// (src/tensor.jl:612-642, :864-899) -- two LAPACK getrs sweeps around two permutedims!.  On B200 the cached
// object is the explicit inverse of each (small, square) factor, formed ONCE here on the device, so that every
// subsequent ldiv! is the same pair of tensor-core mode products as mul! (scimlop_kron_mul on the inverse
// factors: 35 TFLOP/s instead of triangular solves).  Forming the inverse is setup work, like `lu`.
//
//   scimlop_invert   : in-place Gauss-Jordan with partial (row) pivoting, one pivot search + one rank-1 sweep
//                      per column; the matrix (2 MB at 512^2) lives in L2.
//   scimlop_diag_div : W = V ./ d -- Diagonal / BatchedDiagonal ldiv! (src/batch.jl:181-200).
#include <algorithm>

#include "context.h"

namespace {

template <typename T>
__global__ void __launch_bounds__(1024) pivot_kernel(const T* __restrict__ a, long long lda, int n, int k,
                                                      T* __restrict__ f, int* __restrict__ piv, int* info) {
  // argmax_{i >= k} |a[i,k]| (first maximum, like idamax), then f = column k with rows k <-> p exchanged
  __shared__ double smax[32];
  __shared__ int sidx[32];
  __shared__ int sp;
  const T* col = a + (long long)k * lda;
  double best = -1.0;
  int bi = k;
  for (int i = k + threadIdx.x; i < n; i += blockDim.x) {
    const double v = fabs((double)col[i]);
    if (v > best) { best = v; bi = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, best, o);
    const int oi = __shfl_down_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = best; sidx[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = threadIdx.x < (blockDim.x >> 5) ? smax[threadIdx.x] : -1.0;
    bi = threadIdx.x < (blockDim.x >> 5) ? sidx[threadIdx.x] : n;
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, best, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) {
      sp = bi;
      piv[k] = bi;
      if (!(best > 0.0) && *info == 0) *info = k + 1;  // exactly singular (or NaN) at this column
    }
  }
  __syncthreads();
  const int p = sp;
  for (int i = threadIdx.x; i < n; i += blockDim.x) f[i] = col[i == k ? p : (i == p ? k : i)];
}

template <typename T>
__global__ void __launch_bounds__(256) eliminate_kernel(T* __restrict__ a, long long lda, int n, int k,
                                                        const T* __restrict__ f, const int* __restrict__ piv) {
  // one CTA per column j: exchange rows k <-> p, then a[:,j] -= f * (a[k,j] / pivot); column k is the in-place
  // column of the inverse (it starts this step as e_k)
  const int p = piv[k];
  const T pivot = f[k];
  for (int j = blockIdx.x; j < n; j += gridDim.x) {
    T* col = a + (long long)j * lda;
    const bool own = (j == k);
    const T akj = own ? T(1) : col[p];   // row k after the exchange
    const T apj = own ? T(0) : col[k];   // row p after the exchange
    __syncthreads();                     // every thread has read rows k and p before anyone overwrites them
    const T r = akj / pivot;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      T v;
      if (i == k) v = r;
      else {
        const T old = (i == p) ? apj : (own ? T(0) : col[i]);
        v = old - f[i] * r;
      }
      col[i] = v;
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void unpivot_kernel(T* __restrict__ a, long long lda, int n, const int* __restrict__ piv) {
  // inv(A) = inv(P A) P: undo the row exchanges as column exchanges, last first; rows are independent
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = n - 1; k >= 0; k--) {
    const int p = piv[k];
    if (p != k) {
      const T x = a[i + (long long)k * lda];
      a[i + (long long)k * lda] = a[i + (long long)p * lda];
      a[i + (long long)p * lda] = x;
    }
  }
}

template <typename T>
__global__ void diag_div_kernel(long long n, long long k, const T* __restrict__ d, long long ldd,
                                const T* __restrict__ V, long long ldv, T* __restrict__ W, long long ldw) {
  const long long total = n * k;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long j = e / n, i = e - j * n;
    W[i + j * ldw] = V[i + j * ldv] / d[i + j * ldd];
  }
}

inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

template <typename T>
int invert_t(scimlop_context* c, int n, const T* A, long long lda, T* X, long long ldx, void* ws, cudaStream_t st) {
  T* f = static_cast<T*>(ws);
  int* piv = reinterpret_cast<int*>(static_cast<char*>(ws) + align256((size_t)n * sizeof(T)));
  int* info = piv + n;
  if (A != X)
    SCIMLOP_CUDA(c, cudaMemcpy2DAsync(X, (size_t)ldx * sizeof(T), A, (size_t)lda * sizeof(T), (size_t)n * sizeof(T), n,
                                      cudaMemcpyDeviceToDevice, st));
  SCIMLOP_CUDA(c, cudaMemsetAsync(info, 0, sizeof(int), st));
  const int grid = std::min(n, 8 * c->sm_count);
  for (int k = 0; k < n; k++) {
    pivot_kernel<T><<<1, 1024, 0, st>>>(X, ldx, n, k, f, piv, info);
    eliminate_kernel<T><<<grid, 256, 0, st>>>(X, ldx, n, k, f, piv);
  }
  unpivot_kernel<T><<<(n + 127) / 128, 128, 0, st>>>(X, ldx, n, piv);
  c->launches += 2ll * n + 1;
  int hinfo = 0;
  SCIMLOP_CUDA(c, cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, st));
  SCIMLOP_CUDA(c, cudaStreamSynchronize(st));
  SCIMLOP_CUDA(c, cudaGetLastError());
  if (hinfo != 0)
    return scimlop_fail(c, SCIMLOP_ERR_SINGULAR, "invert: zero pivot at column %d (matrix is singular to working precision)", hinfo);
  return SCIMLOP_OK;
}

}  // namespace

extern "C" int scimlop_invert_workspace_bytes(int dtype, int64_t n, size_t* bytes) {
  if (!bytes) return SCIMLOP_ERR_NULL_POINTER;
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32) return SCIMLOP_ERR_UNSUPPORTED;
  if (n < 0 || n > (1 << 20)) return SCIMLOP_ERR_BAD_SHAPE;
  *bytes = align256((size_t)n * (dtype == SCIMLOP_F64 ? 8 : 4)) + align256(((size_t)n + 1) * sizeof(int));
  return SCIMLOP_OK;
}

extern "C" int scimlop_invert(scimlop_handle_t h, int dtype, int64_t n, const void* A, int64_t lda, void* Ainv,
                              int64_t ldinv, void* workspace, size_t workspace_bytes, void* stream) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
  if (n < 0 || n > (1 << 20)) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "invert: bad order");
  if (n == 0) return SCIMLOP_OK;
  if (!A || !Ainv) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "invert: null matrix");
  if (lda < n || ldinv < n) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "invert: leading dimension smaller than n");
  if (A == Ainv && lda != ldinv) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "invert: in-place call with two leading dimensions");
  size_t need = 0;
  scimlop_invert_workspace_bytes(dtype, n, &need);
  if (!workspace || workspace_bytes < need)
    return scimlop_fail(h, SCIMLOP_ERR_WORKSPACE, "invert: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == SCIMLOP_F64
             ? invert_t<double>(h, (int)n, static_cast<const double*>(A), lda, static_cast<double*>(Ainv), ldinv, workspace, st)
             : invert_t<float>(h, (int)n, static_cast<const float*>(A), lda, static_cast<float*>(Ainv), ldinv, workspace, st);
}

extern "C" int scimlop_diag_div(scimlop_handle_t h, int dtype, int64_t n, int64_t k, const void* d, int64_t ldd,
                                const void* V, int64_t ldv, void* W, int64_t ldw, void* stream) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
  if (n < 0 || k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "diag_div: negative size");
  if (n == 0 || k == 0) return SCIMLOP_OK;
  if (!d || !V || !W) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "diag_div: null operand");
  if (ldv < n || ldw < n || (ldd != 0 && ldd < n))
    return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "diag_div: leading dimension smaller than n");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = (long long)n * k;
  const int grid = (int)std::min<long long>((total + 255) / 256, 16ll * h->sm_count);
  if (dtype == SCIMLOP_F64)
    diag_div_kernel<double><<<grid, 256, 0, st>>>(n, k, static_cast<const double*>(d), ldd, static_cast<const double*>(V), ldv,
                                                  static_cast<double*>(W), ldw);
  else
    diag_div_kernel<float><<<grid, 256, 0, st>>>(n, k, static_cast<const float*>(d), ldd, static_cast<const float*>(V), ldv,
                                                 static_cast<float*>(W), ldw);
  h->launches += 1;
  SCIMLOP_CUDA(h, cudaGetLastError());
  return SCIMLOP_OK;
}
