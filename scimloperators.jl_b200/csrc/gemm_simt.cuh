// gemm_simt.cuh -- shape/alignment-agnostic strided-batched GEMM (FP64 DFMA / FP32 FFMA).
//
// Takes every case the TMA kernels cannot: odd leading dimensions, unaligned base pointers, tiny
// factors (the reference's own test shapes are 3x5, 7x11, 13x17 with K = 19, test/matrix.jl:510-517).
// Same contract and fused epilogue as gemm_dmma.cuh.  Plain guarded loads -> padded shared tiles ->
// 4x4 register blocking.  Not a speed-of-light kernel; it is the correctness net under the fast paths
// (there is no CPU fallback anywhere).
#pragma once
#include <cuda_runtime.h>

namespace scimlop {
namespace simt {

constexpr int TM = 64, TN = 64, TK = 16, THREADS = 256;
constexpr int MAX_SCALES = 3;

template <typename T>
struct Scale {
  const T* ptr;
  long long ld;  // 0: vector indexed by row m; else m + n*ld
};

template <typename T>
struct Params {
  int M, N, K, batch;
  int tiles_m, tiles_n;
  const T* A;
  long long lda, strideA;
  const T* B;
  long long ldb, strideB;
  T* C;
  long long ldc, strideC;
  T alpha, beta;  // beta == 0: C never read
  int nscale;
  Scale<T> scale[MAX_SCALES];
};

// A_KC: A(m,k) at k + m*lda (transposed storage), else m + k*lda.
// B_OC: B(k,n) at n + k*ldb (transposed storage), else k + n*ldb.
template <typename T, bool A_KC, bool B_OC>
__global__ void __launch_bounds__(THREADS) gemm_simt_kernel(const Params<T> p) {
  __shared__ T As[TK][TM + 1];
  __shared__ T Bs[TK][TN + 1];

  const long long tile = blockIdx.x;
  const int mt = (int)(tile % p.tiles_m);
  const long long r = tile / p.tiles_m;
  const int nt = (int)(r % p.tiles_n);
  const int b = (int)(r / p.tiles_n);
  const int m0 = mt * TM, n0 = nt * TN;
  const T* __restrict__ A = p.A + (long long)b * p.strideA;
  const T* __restrict__ B = p.B + (long long)b * p.strideB;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = T(0);

  for (int k0 = 0; k0 < p.K; k0 += TK) {
#pragma unroll
    for (int q = 0; q < (TM * TK) / THREADS; q++) {
      const int idx = tid + q * THREADS;
      int m, k;
      if (A_KC) { k = idx % TK; m = idx / TK; } else { m = idx % TM; k = idx / TM; }
      const int gm = m0 + m, gk = k0 + k;
      T v = T(0);
      if (gm < p.M && gk < p.K) v = A_KC ? A[gk + (long long)gm * p.lda] : A[gm + (long long)gk * p.lda];
      As[k][m] = v;
    }
#pragma unroll
    for (int q = 0; q < (TN * TK) / THREADS; q++) {
      const int idx = tid + q * THREADS;
      int n, k;
      if (B_OC) { n = idx % TN; k = idx / TN; } else { k = idx % TK; n = idx / TK; }
      const int gn = n0 + n, gk = k0 + k;
      T v = T(0);
      if (gn < p.N && gk < p.K) v = B_OC ? B[gn + (long long)gk * p.ldb] : B[gk + (long long)gn * p.ldb];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; k++) {
      T a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[k][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; j++) bb[j] = Bs[k][ty + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  T* C = p.C + (long long)b * p.strideC;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int n = n0 + ty + 16 * j;
    if (n >= p.N) continue;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int m = m0 + tx + 16 * i;
      if (m >= p.M) continue;
      T v = acc[i][j];
      for (int q = 0; q < p.nscale; q++)
        v *= p.scale[q].ptr[(p.scale[q].ld ? (long long)n * p.scale[q].ld : 0) + m];
      v *= p.alpha;
      T* dst = C + (long long)n * p.ldc + m;
      if (p.beta != T(0)) v += p.beta * *dst;
      *dst = v;
    }
  }
}

}  // namespace simt
}  // namespace scimlop
