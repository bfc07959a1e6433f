/* ref_c.c -- CPU oracle, C restatement (TEST INFRASTRUCTURE ONLY; see oracle/ref_numpy.py header).
 *
 * Restates, in plain C, the two reference algorithms for the k>1 outer stage of
 * TensorProductOperator `mul!` and the sweeps the operator-tree path performs, so that they can be
 * (1) cross-checked against the NumPy restatement and (2) timed as the CPU baseline ("port").
 *
 *   stdlib path : src/tensor.jl:568 (stage-1 GEMM), :778-788 permutedims!→GEMM→permutedims!,
 *                 :821-831 (+ copy!(c4,w), axpby!(β,c4,α,w))
 *   LV path     : ext/SciMLOperatorsLoopVectorizationExt.jl:17-23 and :35-41 (single-thread SIMD nest)
 *   sweeps      : src/batch.jl:151-179, stdlib Diagonal mul! (src/matrix.jl:340,349 with A::Diagonal),
 *                 src/basic.jl:74-90 (identity copy/axpby), :248-264 (lmul!)
 *
 * The GEMM here is a plain cache-blocked loop nest -- in the reference it is OpenBLAS gemm! (a
 * dependency outside /root/reference); bench.py uses NumPy's OpenBLAS for that stage when timing.
 * All arrays column-major.  Build: see oracle/Makefile (gcc -O3 -march=native, no OpenMP: Base
 * permutedims!/copy!/axpby! and @turbo are single-threaded).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DEFINE_ORACLE(T, S)                                                                         \
/* C(M×N) = α·op(A)·B + β·C, op(A)=A (ta=0, A is M×K) or Aᵀ (ta=1, A is K×M). β==0 ⇒ C not read */ \
void oracle_gemm_##S(int ta, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t lda,     \
                     const T* B, int64_t ldb, T beta, T* C, int64_t ldc) {                          \
  for (int64_t j = 0; j < N; j++) {                                                                 \
    T* c = C + j * ldc;                                                                             \
    if (beta == (T)0) { for (int64_t i = 0; i < M; i++) c[i] = (T)0; }                              \
    else if (beta != (T)1) { for (int64_t i = 0; i < M; i++) c[i] *= beta; }                        \
    if (!ta) {                                                                                      \
      for (int64_t p = 0; p < K; p++) {                                                             \
        const T b = alpha * B[p + j * ldb]; const T* a = A + p * lda;                               \
        for (int64_t i = 0; i < M; i++) c[i] += a[i] * b;                                           \
      }                                                                                             \
    } else {                                                                                        \
      for (int64_t i = 0; i < M; i++) {                                                             \
        const T* a = A + i * lda; const T* b = B + j * ldb; T acc = (T)0;                           \
        for (int64_t p = 0; p < K; p++) acc += a[p] * b[p];                                         \
        c[i] += alpha * acc;                                                                        \
      }                                                                                             \
    }                                                                                               \
  }                                                                                                 \
}                                                                                                   \
/* permutedims!(dst, src, (2,1,3)): src (d1,d2,d3) → dst (d2,d1,d3)  [src/tensor.jl:780,785] */     \
void oracle_permute213_##S(T* dst, const T* src, int64_t d1, int64_t d2, int64_t d3) {              \
  for (int64_t l = 0; l < d3; l++)                                                                  \
    for (int64_t i = 0; i < d1; i++)                                                                \
      for (int64_t j = 0; j < d2; j++)                                                              \
        dst[j + d2 * (i + d1 * l)] = src[i + d1 * (j + d2 * l)];                                    \
}                                                                                                   \
/* LV hook: W[i,m,j] = Σ_o A[m,o]·C[i,o,j]; use_ab ⇒ W = α·acc + β·W  [ext/...LV...:17-23,35-41] */ \
void oracle_lv_outer_mul_##S(T* w, const T* A, int64_t lda, const T* C, int64_t mi, int64_t mo,     \
                             int64_t no, int64_t k, int use_ab, T alpha, T beta) {                  \
  for (int64_t j = 0; j < k; j++)                                                                   \
    for (int64_t m = 0; m < mo; m++) {                                                              \
      T* wc = w + mi * (m + mo * j);                                                                \
      for (int64_t i0 = 0; i0 < mi; i0 += 64) {                                                     \
        int64_t ib = mi - i0 < 64 ? mi - i0 : 64; T acc[64];                                        \
        for (int64_t i = 0; i < ib; i++) acc[i] = (T)0;                                             \
        for (int64_t o = 0; o < no; o++) {                                                          \
          const T a = A[m + o * lda]; const T* c = C + i0 + mi * (o + no * j);                      \
          for (int64_t i = 0; i < ib; i++) acc[i] += a * c[i];                                      \
        }                                                                                           \
        if (!use_ab) for (int64_t i = 0; i < ib; i++) wc[i0 + i] = acc[i];                          \
        else if (beta == (T)0) for (int64_t i = 0; i < ib; i++) wc[i0 + i] = alpha * acc[i];        \
        else for (int64_t i = 0; i < ib; i++) wc[i0 + i] = alpha * acc[i] + beta * wc[i0 + i];      \
      }                                                                                             \
    }                                                                                               \
}                                                                                                   \
/* stdlib outer stage: permute → GEMM → [copy] → permute → [axpby]  [src/tensor.jl:778-788,821-831] \
   scratch: c2 (no·mi·k), c3 (mo·mi·k), c4 (mo·mi·k; only when use_ab && beta!=0) */                \
void oracle_stdlib_outer_mul_##S(T* w, const T* A, int64_t lda, const T* C1, int64_t mi, int64_t mo,\
                                 int64_t no, int64_t k, int use_ab, T alpha, T beta,                \
                                 T* c2, T* c3, T* c4) {                                             \
  oracle_permute213_##S(c2, C1, mi, no, k);                                                         \
  oracle_gemm_##S(0, mo, mi * k, no, (T)1, A, lda, c2, no, (T)0, c3, mo);                           \
  int64_t n = mi * mo * k;                                                                          \
  if (use_ab && beta != (T)0) memcpy(c4, w, (size_t)n * sizeof(T));                                 \
  oracle_permute213_##S(w, c3, mo, mi, k);                                                          \
  if (use_ab) {                                                                                     \
    if (beta == (T)0) for (int64_t i = 0; i < n; i++) w[i] = alpha * w[i];                          \
    else for (int64_t i = 0; i < n; i++) w[i] = alpha * w[i] + beta * c4[i];                        \
  }                                                                                                 \
}                                                                                                   \
/* whole 2-factor `mul!`: w = [α]·(A⊗B)·v [+ β·w]; A outer mo×no, B inner mi×ni, v (ni·no)×k.       \
   lv=1 takes the LV hook for k>1, lv=0 the stdlib path; k==1 is the transposed-GEMM shortcut      \
   (src/tensor.jl:766-770, :809-813).  Returns 0, or -1 if scratch allocation failed. */            \
int oracle_tpo2_mul_##S(T* w, const T* A, const T* B, const T* v, int64_t mo, int64_t no,           \
                        int64_t mi, int64_t ni, int64_t k, int use_ab, T alpha, T beta, int lv) {   \
  T* c1 = (T*)malloc((size_t)(mi * no * k) * sizeof(T));                                            \
  if (!c1) return -1;                                                                               \
  oracle_gemm_##S(0, mi, no * k, ni, (T)1, B, mi, v, ni, (T)0, c1, mi);         /* :568 / :603 */   \
  if (k == 1) {                                                                                     \
    /* W(mi×mo) = α·C1(mi×no)·Aᵀ + β·W : row m of A against columns of C1 */                        \
    for (int64_t m = 0; m < mo; m++)                                                                \
      for (int64_t i = 0; i < mi; i++) {                                                            \
        T acc = (T)0;                                                                               \
        for (int64_t o = 0; o < no; o++) acc += c1[i + mi * o] * A[m + mo * o];                     \
        T* d = w + i + mi * m;                                                                      \
        *d = !use_ab ? acc : (beta == (T)0 ? alpha * acc : alpha * acc + beta * *d);                \
      }                                                                                             \
  } else if (lv) {                                                                                  \
    oracle_lv_outer_mul_##S(w, A, mo, c1, mi, mo, no, k, use_ab, alpha, beta);                      \
  } else {                                                                                          \
    size_t n2 = (size_t)(no * mi * k), n3 = (size_t)(mo * mi * k);                                  \
    T* c2 = (T*)malloc(n2 * sizeof(T)); T* c3 = (T*)malloc(n3 * sizeof(T));                         \
    T* c4 = (T*)malloc(n3 * sizeof(T));                                                             \
    if (!c2 || !c3 || !c4) { free(c1); free(c2); free(c3); free(c4); return -1; }                   \
    oracle_stdlib_outer_mul_##S(w, A, mo, c1, mi, mo, no, k, use_ab, alpha, beta, c2, c3, c4);      \
    free(c2); free(c3); free(c4);                                                                   \
  }                                                                                                 \
  free(c1);                                                                                         \
  return 0;                                                                                         \
}                                                                                                   \
/* w = [α]·(d ⊙ v) [+ β·w]; batched=0: d has n entries broadcast along k columns (Diagonal mul!);    \
   batched=1: d is n×k (src/batch.jl:151-179, :222-226). β==0 ⇒ w not read. */                      \
void oracle_diag_mul_##S(T* w, const T* d, const T* v, int64_t n, int64_t k, int batched,           \
                         int use_ab, T alpha, T beta) {                                             \
  for (int64_t j = 0; j < k; j++) {                                                                 \
    const T* dj = batched ? d + n * j : d; const T* vj = v + n * j; T* wj = w + n * j;              \
    if (!use_ab) for (int64_t i = 0; i < n; i++) wj[i] = dj[i] * vj[i];                             \
    else if (beta == (T)0) for (int64_t i = 0; i < n; i++) wj[i] = alpha * (dj[i] * vj[i]);         \
    else for (int64_t i = 0; i < n; i++) wj[i] = alpha * (dj[i] * vj[i]) + beta * wj[i];            \
  }                                                                                                 \
}                                                                                                   \
/* w = α·v + β·w (identity 5-arg, src/basic.jl:89; axpby!) ; β==0 ⇒ w not read */                   \
void oracle_axpby_##S(T* w, const T* v, int64_t n, T alpha, T beta) {                               \
  if (beta == (T)0) for (int64_t i = 0; i < n; i++) w[i] = alpha * v[i];                            \
  else for (int64_t i = 0; i < n; i++) w[i] = alpha * v[i] + beta * w[i];                           \
}                                                                                                   \
/* lmul!(β, w) with strong zero (src/basic.jl:251,263) */                                           \
void oracle_lmul_##S(T* w, int64_t n, T beta) {                                                     \
  if (beta == (T)0) memset(w, 0, (size_t)n * sizeof(T));                                            \
  else for (int64_t i = 0; i < n; i++) w[i] *= beta;                                                \
}

DEFINE_ORACLE(double, d)
DEFINE_ORACLE(float, s)
