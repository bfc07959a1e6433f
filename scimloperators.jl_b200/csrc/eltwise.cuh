// eltwise.cuh -- the single vectorised pass that evaluates every elementwise part of an operator tree:
//
//   W[i,j] = alpha * ( sum_t coef_t * prod_f d_{t,f}(i,j) ) * V[i,j] + beta * W[i,j]
//
// d_{t,f} is either a vector diagonal (DiagonalOperator, src/matrix.jl:438-459: d[i], or, for a
// diagonal factor inside a Kronecker product, d[(i / div) % mod]) or a batched diagonal
// (BatchedDiagonalOperator, src/batch.jl:151-179: d[i + j*ld]).  A term with no factors is the
// IdentityOperator (src/basic.jl:74-90); zero terms is the NullOperator / lmul!(β, w)
// (src/basic.jl:248-264).  The reference performs one read-modify-write sweep of W per summand
// (src/basic.jl:834-854) plus copy!/lmul!/axpy! sweeps per composition (:1191-1207); here V, W and each
// batched diagonal cross HBM exactly once.  beta == 0 never reads W (strong zero).
//
// HBM-bound: 16-byte streaming loads/stores, 4 independent columns in flight per thread, vector
// diagonals hoisted out of the column loop, grid = row-chunks x column-chunks.
#pragma once
#include <cuda_runtime.h>

namespace scimlop {
namespace elt {

constexpr int MAX_TERMS = 4, MAX_FAC = 4, THREADS = 256, UNROLL = 4;

template <typename T>
struct Factor {
  const T* ptr;
  long long ld;        // != 0: batched, element i + j*ld ; == 0: vector, element (i / div) % mod
  long long div, mod;  // vector factors only
};

template <typename T>
struct Params {
  long long n, k;
  const T* V;
  long long ldv;
  T* W;
  long long ldw;
  T alpha, beta;
  int nterms;
  T coef[MAX_TERMS];
  int nvec[MAX_TERMS];  // factors [0, nvec) are vector diagonals, [nvec, nfac) batched
  int nfac[MAX_TERMS];
  Factor<T> fac[MAX_TERMS][MAX_FAC];
  int cols_per_block;
};

template <typename T, int VEC>
struct Vec;
template <> struct Vec<double, 2> { using type = double2; };
template <> struct Vec<float, 4> { using type = float4; };
template <> struct Vec<double, 1> { using type = double; };
template <> struct Vec<float, 1> { using type = float; };

template <typename T, int VEC>
__device__ __forceinline__ void ld_stream(T (&dst)[VEC], const T* p) {
  using V = typename Vec<T, VEC>::type;
  const V v = __ldcs(reinterpret_cast<const V*>(p));
  const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
  for (int q = 0; q < VEC; q++) dst[q] = e[q];
}
template <typename T, int VEC>
__device__ __forceinline__ void st_stream(T* p, const T (&src)[VEC]) {
  using V = typename Vec<T, VEC>::type;
  V v;
  T* e = reinterpret_cast<T*>(&v);
#pragma unroll
  for (int q = 0; q < VEC; q++) e[q] = src[q];
  __stcs(reinterpret_cast<V*>(p), v);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(THREADS) eltwise_kernel(const Params<T> p) {
  const long long i0 = ((long long)blockIdx.x * THREADS + threadIdx.x) * VEC;
  if (i0 >= p.n) return;
  const long long j_begin = (long long)blockIdx.y * p.cols_per_block;
  const long long j_end = min(p.k, j_begin + (long long)p.cols_per_block);

  // column-independent part of every term
  T rc[MAX_TERMS][VEC];
#pragma unroll
  for (int t = 0; t < MAX_TERMS; t++) {
#pragma unroll
    for (int q = 0; q < VEC; q++) rc[t][q] = T(0);
    if (t < p.nterms) {
#pragma unroll
      for (int q = 0; q < VEC; q++) rc[t][q] = p.coef[t];
      for (int f = 0; f < p.nvec[t]; f++) {
        const Factor<T> F = p.fac[t][f];
#pragma unroll
        for (int q = 0; q < VEC; q++) rc[t][q] *= __ldg(F.ptr + ((i0 + q) / F.div) % F.mod);
      }
    }
  }
  // fold terms without batched factors into one coefficient
  T base[VEC];
#pragma unroll
  for (int q = 0; q < VEC; q++) base[q] = T(0);
  bool any_batched = false;
#pragma unroll
  for (int t = 0; t < MAX_TERMS; t++) {
    if (t < p.nterms) {
      if (p.nfac[t] == p.nvec[t]) {
#pragma unroll
        for (int q = 0; q < VEC; q++) base[q] += rc[t][q];
      } else {
        any_batched = true;
      }
    }
  }
  const bool readw = (p.beta != T(0));
  const bool readv = (p.nterms > 0);

  for (long long j = j_begin; j < j_end; j += UNROLL) {
    T v[UNROLL][VEC], w[UNROLL][VEC], e[UNROLL][VEC];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
#pragma unroll
      for (int q = 0; q < VEC; q++) { v[u][q] = T(0); w[u][q] = T(0); e[u][q] = base[q]; }
      if (j + u < j_end) {
        if (readv) ld_stream<T, VEC>(v[u], p.V + (j + u) * p.ldv + i0);
        if (readw) ld_stream<T, VEC>(w[u], p.W + (j + u) * p.ldw + i0);
      }
    }
    if (any_batched) {
#pragma unroll
      for (int t = 0; t < MAX_TERMS; t++) {
        if (t < p.nterms && p.nfac[t] > p.nvec[t]) {
          T pr[UNROLL][VEC];
#pragma unroll
          for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int q = 0; q < VEC; q++) pr[u][q] = rc[t][q];
          for (int f = p.nvec[t]; f < p.nfac[t]; f++) {
            const Factor<T> F = p.fac[t][f];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
              if (j + u < j_end) {
                T d[VEC];
                ld_stream<T, VEC>(d, F.ptr + (j + u) * F.ld + i0);
#pragma unroll
                for (int q = 0; q < VEC; q++) pr[u][q] *= d[q];
              }
            }
          }
#pragma unroll
          for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int q = 0; q < VEC; q++) e[u][q] += pr[u][q];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (j + u < j_end) {
        T o[VEC];
#pragma unroll
        for (int q = 0; q < VEC; q++) {
          T x = readv ? p.alpha * (e[u][q] * v[u][q]) : T(0);
          if (readw) x += p.beta * w[u][q];
          o[q] = x;
        }
        st_stream<T, VEC>(p.W + (j + u) * p.ldw + i0, o);
      }
    }
  }
}

// Specialisation for programs without batched diagonals (Diagonal / Identity / Null / axpby and any sum of
// products of vector diagonals): the whole tree collapses to one per-row coefficient, so the kernel streams
// only V (and W when beta != 0).  Few registers -> 8 columns in flight per thread at full occupancy, which is
// what a pure-streaming pass needs to approach the HBM roofline.
template <typename T, int VEC>
__global__ void __launch_bounds__(THREADS) eltwise_rowcoef_kernel(const Params<T> p) {
  constexpr int U = 8;
  const long long i0 = ((long long)blockIdx.x * THREADS + threadIdx.x) * VEC;
  if (i0 >= p.n) return;
  const long long j_begin = (long long)blockIdx.y * p.cols_per_block;
  const long long j_end = min(p.k, j_begin + (long long)p.cols_per_block);
  T base[VEC];
#pragma unroll
  for (int q = 0; q < VEC; q++) base[q] = T(0);
  for (int t = 0; t < p.nterms; t++) {
    T c[VEC];
#pragma unroll
    for (int q = 0; q < VEC; q++) c[q] = p.coef[t];
    for (int f = 0; f < p.nfac[t]; f++) {
      const Factor<T> F = p.fac[t][f];
#pragma unroll
      for (int q = 0; q < VEC; q++) c[q] *= __ldg(F.ptr + ((i0 + q) / F.div) % F.mod);
    }
#pragma unroll
    for (int q = 0; q < VEC; q++) base[q] += c[q];
  }
#pragma unroll
  for (int q = 0; q < VEC; q++) base[q] *= p.alpha;
  const bool readw = (p.beta != T(0));
  const bool readv = (p.nterms > 0);
  const T* vp = p.V + j_begin * p.ldv + i0;
  T* wp = p.W + j_begin * p.ldw + i0;
  for (long long j = j_begin; j < j_end; j += U, vp += U * p.ldv, wp += U * p.ldw) {
    T v[U][VEC], w[U][VEC];
#pragma unroll
    for (int u = 0; u < U; u++) {
#pragma unroll
      for (int q = 0; q < VEC; q++) { v[u][q] = T(0); w[u][q] = T(0); }
      if (j + u < j_end) {
        if (readv) ld_stream<T, VEC>(v[u], vp + u * p.ldv);
        if (readw) ld_stream<T, VEC>(w[u], wp + u * p.ldw);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (j + u < j_end) {
        T o[VEC];
#pragma unroll
        for (int q = 0; q < VEC; q++) {
          T x = readv ? base[q] * v[u][q] : T(0);
          if (readw) x += p.beta * w[u][q];
          o[q] = x;
        }
        st_stream<T, VEC>(wp + u * p.ldw, o);
      }
    }
  }
}

}  // namespace elt
}  // namespace scimlop
