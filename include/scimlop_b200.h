/* scimlop_b200.h -- C ABI of libscimlop_b200.so, the B200-native evaluation backend for the cached
 * in-place `mul!` path of SciMLOperators.jl (TensorProductOperator and Added/Composed/Scaled trees
 * over Matrix/Diagonal/BatchedDiagonal/Identity/Null operators).
 *
 * Conventions:
 *   - every array pointer is a DEVICE pointer unless the function name ends in `_host`;
 *   - all matrices are column-major with explicit leading dimensions (in elements);
 *   - `alpha`/`beta` are HOST pointers to one scalar of the operand dtype; `beta == NULL` or `*beta == 0`
 *     means "3-arg mul!": W is overwritten and never read (src/batch.jl:222-223, src/basic.jl:251,521,
 *     BLAS convention at src/matrix.jl:349); `alpha == NULL` means 1;
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*).  A `mul!` on a CACHED operator
 *     never allocates and never re-prepares constant data: scratch is caller-owned (`*_workspace_bytes` says how
 *     much) and the hi / lo images that the Float32 tensor-core GEMM needs of a large constant matrix are made once
 *     (scimlop_split_attach, below) -- mirrors `cache_operator` allocating once and `mul!` being allocation-free
 *     (test/Downstream/alloccheck.jl).  Only the two convenience forms say otherwise in their own comments:
 *     scimlop_kron_mul_host (device staging) and a large Float32 product on an operand that was never attached
 *     (handle-owned grow-only scratch + one extra split launch per call);
 *   - every function returns an int status (SCIMLOP_OK == 0); no exception crosses the boundary;
 *   - a handle is per device and is not thread-safe, like a cached operator's scratch
 *     (src/SciMLOperators.jl:131-132).
 *
 * The reference is pure Julia and has no FFI of its own; each entry point below names the reference
 * method whose body it replaces (file:line under the SciMLOperators.jl tree).  INTEGRATION.md shows the
 * `ccall` stubs of the package extension that binds them.
 */
#ifndef SCIMLOP_B200_H
#define SCIMLOP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCIMLOP_VERSION 100 /* 0.1.0 */

/* ---- status codes ------------------------------------------------------------------------------ */
enum {
  SCIMLOP_OK = 0,
  SCIMLOP_ERR_NULL_POINTER = 1,  /* a required pointer is NULL */
  SCIMLOP_ERR_BAD_SHAPE = 2,     /* negative size, ld < rows, inconsistent factor dims (the @assert's) */
  SCIMLOP_ERR_WORKSPACE = 3,     /* workspace missing or too small (`@assert iscached(L)`, tensor.jl:548) */
  SCIMLOP_ERR_UNSUPPORTED = 4,   /* dtype / precision / factor kind not supported */
  SCIMLOP_ERR_CUDA = 5,          /* a CUDA runtime/driver call failed; see scimlop_last_error */
  SCIMLOP_ERR_NCCL = 6,          /* NCCL missing or a collective failed */
  SCIMLOP_ERR_BAD_HANDLE = 7,
  SCIMLOP_ERR_NO_DEVICE = 8,     /* no usable sm_100 device: there is NO CPU fallback */
  SCIMLOP_ERR_SINGULAR = 9       /* scimlop_invert met an exactly zero pivot (LAPACK's info > 0 / SingularException) */
};

/* ---- element types & precision modes ------------------------------------------------------------- */
enum { SCIMLOP_F64 = 0, SCIMLOP_F32 = 1,
       /* ComplexF32 / ComplexF64, interleaved (re, im) as Julia stores them; alpha / beta point to (re, im) pairs; leading
        * dimensions count complex elements.  Taken by scimlop_kron_workspace_bytes / scimlop_kron_mul with dense
        * (SCIMLOP_DENSE, | SCIMLOP_TRANS = the lazy adjoint F^H) and identity factors -- one factor is a complex
        * MatrixOperator mul! (test/basic.jl:405-426, test/Downstream/alloccheck.jl:37) -- and by scimlop_csr_mul (sparse
        * complex operators, the direct gather / scatter kernels).  Every other entry answers
        * SCIMLOP_ERR_UNSUPPORTED for them. */
       SCIMLOP_C64 = 2, SCIMLOP_C128 = 3 };
enum {
  SCIMLOP_PREC_DEFAULT = 0,   /* F64: DMMA (mma.sync f64) tensor path.  F32: the fastest path that holds the
                                 1e-5 relative-error contract: 3xTF32 split on tcgen05 tensor cores where the
                                 shape is eligible, exact-FP32 FFMA otherwise                               */
  SCIMLOP_PREC_TF32X3 = 1,    /* F32 only: 3xTF32 split on tcgen05 (same as DEFAULT where eligible)          */
  SCIMLOP_PREC_TF32 = 2,      /* F32 only, explicit opt-in: single-pass TF32 on tcgen05 (rel. err ~1e-3)     */
  SCIMLOP_PREC_FP32_EXACT = 3,/* F32 only: force the exact-FP32 FFMA path                                    */
  SCIMLOP_PREC_BF16 = 4       /* F32 only, explicit opt-in: operands rounded to BF16, one tcgen05.mma kind::f16 pass,
                                 FP32 accumulation (rel. err ~4e-3).  Taken by the fused two-mode Kronecker kernel
                                 (two adjacent dense factors <= 128); every other Float32 shape runs the single-pass
                                 TF32 path, which is at least as accurate                                    */
};

/* ---- factor kinds (kron factors and tree leaves) --------------------------------------------------- */
enum {
  SCIMLOP_DENSE = 0,    /* MatrixOperator over a dense matrix            src/matrix.jl:116-133        */
  SCIMLOP_IDENTITY = 1, /* IdentityOperator                              src/basic.jl:28-30           */
  SCIMLOP_DIAG = 2,     /* DiagonalOperator(::Vector): data = n entries  src/matrix.jl:438-459        */
  SCIMLOP_BDIAG = 3,    /* BatchedDiagonalOperator: data = n x k, ld     src/batch.jl:11-28           */
  SCIMLOP_NULL = 4,     /* NullOperator                                  src/basic.jl:185-188         */
  SCIMLOP_BANDED = 5,   /* MatrixOperator over a banded (tridiagonal, ...) n x n matrix, kron factors only:
                           data = n x (2b+1) column-major, diagonal d in column d+b aligned by row
                           (data[i + n*(d+b)] = A[i, i+d]); lds[f] carries the half-bandwidth b (<= 4)  */
  SCIMLOP_CSR = 6       /* MatrixOperator over a general sparse m x n matrix (CSR), kron factors of scimlop_kron_mul only:
                           factors[f] points to a HOST scimlop_csr_factor describing device arrays; Float32 / Float64,
                           not combinable with SCIMLOP_TRANS (convert once: a CSC matrix is the CSR form of its transpose) */
};
/* A sparse Kronecker factor (`TensorProductOperator(sprand(...), B)`): device pointers, read on the host at call time. */
typedef struct {
  const int32_t* rowptr;   /* rows + 1 entries */
  const int32_t* colidx;   /* nnz entries */
  const void* values;      /* nnz entries of the call's dtype */
  int32_t index_base;      /* 0, or 1 for CUDA.jl's CuSparseMatrixCSR */
  int32_t reserved;
} scimlop_csr_factor;
/* OR-ed into a kron factor kind: the factor is the lazy adjoint/transpose of the stored matrix
 * (`adjoint(::MatrixOperator)` wraps the array, src/matrix.jl:174-175; TPO adjoint maps over factors,
 * src/tensor.jl:181-191).  dims[] are those of the OPERATOR; the stored matrix is cols x rows. */
#define SCIMLOP_TRANS 0x100

typedef struct scimlop_context* scimlop_handle_t;

/* ---- handle ------------------------------------------------------------------------------------------ */
int scimlop_version(void);
const char* scimlop_status_string(int status);
/* Binds the handle to CUDA device `device`; fails with SCIMLOP_ERR_NO_DEVICE unless it is sm_100. */
int scimlop_create(scimlop_handle_t* out, int device);
int scimlop_destroy(scimlop_handle_t h);
/* Copies the text of the last failure recorded on `h` (empty string if none). */
int scimlop_last_error(scimlop_handle_t h, char* buf, size_t len);
/* Explicit behaviour switches of a handle (never environment variables).
 * SCIMLOP_OPT_TAIL_SPLIT (default 1): a Float64 GEMM whose tile count is a few whole waves plus a partial one runs the
 * partial wave as a second, split-K launch (thread-block clusters divide K and reduce through distributed shared
 * memory).  Faster (config 2 on 32 columns: 512 tiles = 3.46 waves -> 3 + 0.5), but the tiles of that wave are summed
 * in a different order than the single-launch schedule; 0 restores one launch and one summation order everywhere
 * (what the fused compute + all-gather entries always use, so a bit-for-bit comparison against them sets 0).
 * SCIMLOP_OPT_PEER_PIPELINE (0 = off, 1 = automatic (default), 2 = always): scimlop_kron_mul_peers / _mc with two dense
 * Float64 factors may run as a pipeline of column chunks (2 unless SCIMLOP_OPT_PEER_CHUNKS says otherwise) on two streams -- the last stage of a chunk, which is
 * bound by NVLink ingress when many ranks gather, on a share of the SMs, the first stage of the next chunk on the rest.
 * Automatic takes it where the estimated gather time is at least 1.25x the last stage's arithmetic (4 and 8 ranks on
 * config 2, not 2 ranks).  Same numbers bit for bit as the unchunked schedule.
 * SCIMLOP_OPT_PEER_CHUNKS (0 = chosen by the library; 2, 4 or 8) and SCIMLOP_OPT_PEER_STAGE2_SMS (0 = chosen; else the
 * number of SMs the last stage of a non-final chunk runs on) fix the shape of that pipeline: tuning knobs, measured by
 * tools/bench_peer_pipeline.py.
 * SCIMLOP_OPT_MODE_PIPELINE (0 = off (default), 2 .. 8 = column chunks): scimlop_kron_mul on three small dense Float32
 * factors runs its two passes (fused two-mode pass, third-mode pass) as a pipeline over that many column chunks on two
 * streams with an SM share each.  Same numbers as the unchunked call. */
enum { SCIMLOP_OPT_TAIL_SPLIT = 1, SCIMLOP_OPT_PEER_PIPELINE = 2, SCIMLOP_OPT_PEER_CHUNKS = 3, SCIMLOP_OPT_PEER_STAGE2_SMS = 4,
       SCIMLOP_OPT_MODE_PIPELINE = 5 };
int scimlop_set_option(scimlop_handle_t h, int option, int value);
/* Number of kernels this handle has launched since creation (bench.py's `gpu_launches`). */
int scimlop_launch_count(scimlop_handle_t h, int64_t* count);

/* ---- leaves ---------------------------------------------------------------------------------------- */
/* MatrixOperator mul!: W(m x k) = alpha * op(A) * V(n x k) + beta * W.    Replaces src/matrix.jl:337-350
 * (`mul!(w, L.A, v[, α, β])` -> BLAS gemm!); transA != 0 is the lazy Adjoint/Transpose wrapper that
 * `adjoint(::MatrixOperator)` of a constant operator builds (src/matrix.jl:174-175): op(A) = A^T with A
 * stored n x m. */
int scimlop_matmul(scimlop_handle_t h, int dtype, int precision, int transA, int64_t m, int64_t n,
                   int64_t k, const void* A, int64_t lda, const void* V, int64_t ldv, void* W,
                   int64_t ldw, const void* alpha, const void* beta, void* stream);

/* MatrixOperator over a general sparse matrix, CSR storage (rowptr: rows + 1 entries, colidx / values: nnz entries,
 * 32-bit indices, `index_base` 0 or 1 -- CUDA.jl's CuSparseMatrixCSR is 1-based):
 *     W = alpha * op(A) * V + beta * W,   op(A) = A (trans == 0: W is rows x k) or A^T (trans != 0: W is cols x k).
 * Replaces `mul!(w, L.A, v[, α, β])` (src/matrix.jl:337-350) for sparse L.A -- the `sprand` operators of
 * test/basic.jl:405-426, test/Downstream/alloccheck.jl:37-45; ext/SciMLOperatorsSparseArraysExt.jl:8-23.  A
 * SparseMatrixCSC is the CSR form of its transpose: bind it with trans flipped.  trans == 0 gathers per output row
 * (deterministic, W never read when beta == 0); trans != 0 scatters with atomic adds (two launches; the summation
 * order, hence the last bits, may differ between runs).  `workspace` (scimlop_csr_workspace_bytes; `cache_operator`
 * allocates it, NULL is allowed) holds a row-contiguous copy of V: with it a gathered row costs full 32-byte sectors
 * instead of one word per sector and column (2.4x at 1024 Float64 columns). */
int scimlop_csr_workspace_bytes(int dtype, int trans, int64_t rows, int64_t cols, int64_t k, size_t* bytes);
int scimlop_csr_mul(scimlop_handle_t h, int dtype, int trans, int64_t rows, int64_t cols, int64_t k,
                    const int32_t* rowptr, const int32_t* colidx, const void* values, int index_base, const void* V,
                    int64_t ldv, void* W, int64_t ldw, const void* alpha, const void* beta, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Float32 operands of the streamed tensor-core GEMM (matrices beyond the 128 x 128 resident-factor kernel: large
 * MatrixOperators, Kronecker factors > 128, dense tree / block-diagonal leaves).  3xTF32 needs the operand as
 * hi + lo halves; for a constant operator that is prepared ONCE: `cache_operator` allocates scimlop_split_bytes
 * and calls scimlop_split_attach (one launch: splits now, and remembers A -> image in the handle);
 * `update_coefficients!` of a time-dependent matrix calls scimlop_split_refresh; the finaliser calls
 * scimlop_split_detach.  Every later product whose operand is (A, lda, rows, cols) -- through scimlop_matmul,
 * scimlop_gemm_strided_batched, scimlop_kron_mul, scimlop_tree_mul, scimlop_blockdiag_mul -- is then ONE kernel
 * launch with no allocation.  rows x cols is the STORED shape (before any transpose flag).  An operand that was
 * never attached still works (uncached convenience): the smaller operand is split per call into a handle-owned
 * scratch that grows on demand.  Replaces nothing in the reference (BLAS sgemm has no such state); it is what keeps
 * `mul!` allocation-free here (test/Downstream/alloccheck.jl:14,100-122). */
int scimlop_split_bytes(int64_t rows, int64_t cols, size_t* bytes);
int scimlop_split_attach(scimlop_handle_t h, const void* A, int64_t lda, int64_t rows, int64_t cols, void* split,
                         size_t split_bytes, void* stream);
int scimlop_split_refresh(scimlop_handle_t h, const void* A, void* stream);
int scimlop_split_detach(scimlop_handle_t h, const void* A);

/* Strided-batched GEMM engine: C_b(M x N) = alpha * op(A_b)(M x K) * op(B_b)(K x N) + beta * C_b,
 * b = 0..batch-1, operand b at base + b*stride (elements; stride 0 = shared operand).  This is what the
 * extension binds for `mul!(::DevMat, ::DevMat, ::DevMat, α, β)` and the transposed-destination form
 * src/matrix.jl:354-361. */
int scimlop_gemm_strided_batched(scimlop_handle_t h, int dtype, int precision, int transA, int transB,
                                 int64_t M, int64_t N, int64_t K, const void* alpha, const void* A,
                                 int64_t lda, int64_t strideA, const void* B, int64_t ldb,
                                 int64_t strideB, const void* beta, void* C, int64_t ldc,
                                 int64_t strideC, int64_t batch, void* stream);

/* Diagonal / batched-diagonal mul!: W(n x k) = alpha * (d .* V) + beta * W.  ldd == 0: d has n entries
 * broadcast along the k columns (stdlib Diagonal mul!, src/matrix.jl:340,349 with A::Diagonal);
 * ldd >= n: d is n x k (src/batch.jl:151-179, call forms :203-228). */
int scimlop_diag_mul(scimlop_handle_t h, int dtype, int64_t n, int64_t k, const void* d, int64_t ldd,
                     const void* V, int64_t ldv, void* W, int64_t ldw, const void* alpha,
                     const void* beta, void* stream);

/* Identity / Null mul!: W(n x k) = alpha * V + beta * W   (src/basic.jl:74-90); V == NULL gives the
 * NullOperator form W = beta * W with a strong zero for beta == 0 (src/basic.jl:248-264). */
int scimlop_axpby(scimlop_handle_t h, int dtype, int64_t n, int64_t k, const void* V, int64_t ldv,
                  void* W, int64_t ldw, const void* alpha, const void* beta, void* stream);

/* ---- ldiv! (the solve side) --------------------------------------------------- */
/* Diagonal / batched-diagonal ldiv!: W(n x k) = V ./ d, ldd as for scimlop_diag_mul; W may alias V (the
 * two-argument form).  src/batch.jl:181-200; stdlib `ldiv!(::Diagonal, v)` behind src/matrix.jl:363-366. */
int scimlop_diag_div(scimlop_handle_t h, int dtype, int64_t n, int64_t k, const void* d, int64_t ldd,
                     const void* V, int64_t ldv, void* W, int64_t ldw, void* stream);

/* The cached object behind `ldiv!` of a square dense factor: Ainv = inv(A), formed on the device by
 * Gauss-Jordan elimination with partial pivoting (Ainv may alias A).  It replaces the LAPACK factorisation of
 * `factorize(L)` (src/matrix.jl:513-516; per Kronecker factor: src/tensor.jl:227): with the inverse factors
 * cached, `ldiv!(w, L::TensorProductOperator, v)` (src/tensor.jl:612-642 + outer_div! :864-899) is
 * scimlop_kron_mul on them -- the same tensor-core mode products as mul!.  Setup work: synchronises `stream`
 * and returns SCIMLOP_ERR_SINGULAR for an exactly zero pivot.  Workspace size from the query. */
int scimlop_invert_workspace_bytes(int dtype, int64_t n, size_t* bytes);
int scimlop_invert(scimlop_handle_t h, int dtype, int64_t n, const void* A, int64_t lda, void* Ainv,
                   int64_t ldinv, void* workspace, size_t workspace_bytes, void* stream);

/* The factorised solve (Float64).  `factorize(L::MatrixOperator)` -> `lu` (src/matrix.jl:513-516; per Kronecker
 * factor src/tensor.jl:227) and `ldiv!` with it (src/tensor.jl:612-673, outer_div! :864-899, LAPACK getrs):
 *
 * scimlop_factorize: blocked LU with partial pivoting on the device (P A = L U; 4 launches per 32 columns: a multi-CTA
 *   panel kernel with the panel rows in registers, row swaps + block-row solve, DMMA trailing update), then the solve
 *   structures (inverted 128 x 128 diagonal blocks, pre-multiplied off-diagonal block rows) into `fact` -- an opaque
 *   caller-owned device blob of scimlop_factorize_bytes -- and, when Ainv != NULL, the explicit inverse (the solve
 *   applied to the identity).  *cond1 receives cond_1(A) = |A|_1 |inv(A)|_1.  Synchronises `stream`; an exactly zero pivot
 *   returns SCIMLOP_ERR_SINGULAR (LAPACK info > 0).  The caller picks the solve per operator: multiplication by Ainv
 *   (scimlop_kron_mul: one GEMM per mode, residual ~ cond * eps -- inside the 1e-12 budget for well-conditioned factors)
 *   or scimlop_kron_ldiv (substitution: residual ~ eps |A| |x| at any cond, 1.25x the flops at n = 512).
 * scimlop_kron_ldiv: W = (F_1 (x) ... (x) F_n)^{-1} V, factor f given by its blob facts[f] (kinds[f] = SCIMLOP_DENSE, or
 *   SCIMLOP_DENSE | SCIMLOP_TRANS to solve with the TRANSPOSE of the factorised matrix -- the lazy adjoint of
 *   src/matrix.jl:565: U^T and L^T sweeps, then P^T) or an identity (SCIMLOP_IDENTITY, facts[f] ignored); dims[f] = order of factor f, outermost first; V, W contiguous
 *   prod(dims) x k and distinct; workspace = prod(dims) * k * 8 bytes.  Every triangular sweep is a sequence of GEMMs
 *   along the factor's mode (no permutedims), out of place between W and the workspace. */
int scimlop_factorize_bytes(int dtype, int64_t n, size_t* fact_bytes, size_t* scratch_bytes);
int scimlop_factorize(scimlop_handle_t h, int dtype, int64_t n, const void* A, int64_t lda, void* fact, size_t fact_bytes,
                      void* Ainv, int64_t ldinv, void* scratch, size_t scratch_bytes, double* cond1, void* stream);
int scimlop_kron_ldiv(scimlop_handle_t h, int dtype, int nfactors, void* const* facts, const int64_t* dims,
                      const int32_t* kinds, const void* V, void* W, int64_t k, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ---- TensorProductOperator ------------------------------------------------------------------------- */
/* Scratch needed by scimlop_kron_mul for these factors and k columns (replaces the 7-slot cache tuple of
 * `cache_self`, src/tensor.jl:460-518: at most two ping-pong buffers, none when <= 1 factor is
 * non-identity).  dims[2*f] = rows, dims[2*f+1] = cols of factor f, outermost factor first. */
int scimlop_kron_workspace_bytes(int dtype, int nfactors, const int64_t* dims, const int32_t* kinds,
                                 int64_t k, size_t* bytes);

/* W(prod rows x k) = alpha * (F_1 (x) F_2 (x) ... (x) F_n) * V(prod cols x k) + beta * W.
 * Replaces `mul!(w, L::TensorProductOperator, v[, α, β])` src/tensor.jl:543-574 / :576-610 including
 * `outer_mul!` :750-834 and the right-fold nesting of the N-ary constructor (:121).  factors[f] is a
 * dense rows x cols matrix (lds[f] >= rows), a length-rows diagonal (SCIMLOP_DIAG) or ignored
 * (SCIMLOP_IDENTITY).  1 <= nfactors <= 8. */
int scimlop_kron_mul(scimlop_handle_t h, int dtype, int precision, int nfactors,
                     const void* const* factors, const int64_t* dims, const int64_t* lds,
                     const int32_t* kinds, const void* V, int64_t ldv, void* W, int64_t ldw, int64_t k,
                     const void* alpha, const void* beta, void* workspace, size_t workspace_bytes,
                     void* stream);

/* The reference's own plugin hook, same contract: W[i,m,j] = sum_o A[m,o] * C[i,o,j]  (then
 * alpha*acc + beta*W[i,m,j]); W is (mi*mo) x k, C is (mi*no) x k, A is mo x no.
 * Replaces `tensor_outer_mul_fast!(w, outer, C, mi, mo, no, k[, α, β])` src/tensor.jl:719, implemented
 * today by ext/SciMLOperatorsLoopVectorizationExt.jl:10-44. */
int scimlop_outer_mul_fast(scimlop_handle_t h, int dtype, int precision, void* W, const void* A,
                           int64_t lda, const void* C, int64_t mi, int64_t mo, int64_t no, int64_t k,
                           const void* alpha, const void* beta, void* stream);

/* TensorSumOperator / kron(A,I) + kron(I,B):  W_j(mi x mo) = alpha*(B*V_j + V_j*A^T) + beta*W_j for
 * every column j, V_j = reshape(V[:,j], mi, mo); A is mo x mo (outer), B is mi x mi (inner).
 * Replaces src/tensor.jl:387-403 (and the AddedOperator of two TPOs of config 5). No workspace. */
int scimlop_kronsum_mul(scimlop_handle_t h, int dtype, int precision, const void* A, int64_t lda,
                        int64_t mo, const void* B, int64_t ldb, int64_t mi, const void* V, int64_t ldv,
                        void* W, int64_t ldw, int64_t k, const void* alpha, const void* beta,
                        void* stream);

/* Kronecker sum of two BANDED matrices (finite-difference Laplacian, the HBM-bound flavour of config 5):
 *   W_j(mi x mo) = alpha * (By * V_j + V_j * Bx^T) + beta * W_j  in ONE pass over V and W.
 * Bx: mo x (2bx+1) bands of the outer factor, By: mi x (2by+1) bands of the inner factor (SCIMLOP_BANDED layout). */
int scimlop_kronsum_banded_mul(scimlop_handle_t h, int dtype, const void* Bx, int bx, int64_t mo, const void* By,
                               int by, int64_t mi, const void* V, int64_t ldv, void* W, int64_t ldw, int64_t k,
                               const void* alpha, const void* beta, void* stream);

/* ---- operator trees --------------------------------------------------------------------------------- */
/* A tree is handed over flattened (the host does what src/basic.jl:619-635 does for sums and :936-939
 * for products): a sum of terms, each term = scale * f_1 * f_2 * ... * f_n (applied right to left). */
typedef struct {
  int32_t kind;     /* SCIMLOP_DENSE / IDENTITY / DIAG / BDIAG / NULL */
  int32_t trans;    /* DENSE only: apply data^T (data stored cols x rows)  */
  const void* data; /* device pointer (NULL for IDENTITY / NULL)          */
  int64_t ld;       /* DENSE: leading dim; BDIAG: leading dim of the n x k diagonal; else ignored */
  int64_t rows;     /* operator rows    */
  int64_t cols;     /* operator columns */
} scimlop_factor_t;

typedef struct {
  double scale; /* product of the ScaledOperator λ's above this term (src/basic.jl:518-536) */
  int32_t nfactors;
  const scimlop_factor_t* factors; /* host array, leftmost factor first */
} scimlop_term_t;

int scimlop_tree_workspace_bytes(int dtype, const scimlop_term_t* terms, int nterms, int64_t k,
                                 size_t* bytes);

/* W(rows x k) = alpha * (sum_t scale_t * prod_f factor_{t,f}) * V + beta * W.
 * Replaces `mul!(w, ::AddedOperator, v[, α, β])` src/basic.jl:834-854, `::ComposedOperator`
 * :1160-1207, `::ScaledOperator` :518-536 and the leaves they recurse into, in one pass over W for the
 * elementwise terms plus one fused-epilogue GEMM per dense term. */
int scimlop_tree_mul(scimlop_handle_t h, int dtype, int precision, const scimlop_term_t* terms,
                     int nterms, const void* V, int64_t ldv, void* W, int64_t ldw, int64_t k,
                     const void* alpha, const void* beta, void* workspace, size_t workspace_bytes,
                     void* stream);

/* BlockDiagonalOperator mul! (src/block.jl:145-179): block b (rows_b x cols_b; DENSE with .trans, DIAG, BDIAG with
 * .ld, IDENTITY or NULL) maps rows [sum cols_<b, +cols_b) of V to rows [sum rows_<b, +rows_b) of W,
 * W_b = alpha * op_b * V_b + beta * W_b.  Equal dense blocks stored at a uniform stride run as ONE strided-batched
 * GEMM (row slices of the column-major V / W are batch strides); otherwise one launch per block. */
int scimlop_blockdiag_mul(scimlop_handle_t h, int dtype, int precision, int nblocks,
                          const scimlop_factor_t* blocks, const void* V, int64_t ldv, void* W, int64_t ldw,
                          int64_t k, const void* alpha, const void* beta, void* stream);

/* ---- host-buffer entry (end-to-end path) ------------------------------------------------------------- */
/* Same as scimlop_kron_mul but V, W and the factors live in (ideally pinned) HOST memory: columns are
 * streamed through the device in chunks of `chunk_cols` (0 = library default) with H2D copy, compute and
 * D2H copy overlapped on three streams.  Device staging is owned by the handle and grown on demand
 * (first call allocates; later calls of the same shape do not).  Blocks until W is complete. */
int scimlop_kron_mul_host(scimlop_handle_t h, int dtype, int precision, int nfactors,
                          const void* const* factors, const int64_t* dims, const int64_t* lds,
                          const int32_t* kinds, const void* V, int64_t ldv, void* W, int64_t ldw,
                          int64_t k, const void* alpha, const void* beta, int64_t chunk_cols);

/* ---- solver-loop vector kernels (the caller side of the matvec) ------------------- */
/* Column-wise dot products: out[j] = sum_i X[i,j] * Y[i,j], j < k; `out` is a DEVICE vector of k scalars.
 * (Y == X gives squared column norms.)  Everything stays on the stream: graph-capturable. */
int scimlop_col_dot(scimlop_handle_t h, int dtype, int64_t n, int64_t k, const void* X, int64_t ldx,
                    const void* Y, int64_t ldy, void* out_dev, void* stream);

/* W[:,j] = a_j * X[:,j] + b_j * W[:,j] with per-column coefficients held in DEVICE memory:
 * a_j = sa * an[j] / ad[j], b_j = sb * bn[j] / bd[j] (an, ad, bn, bd: device vectors of k scalars or NULL = 1;
 * sa, sb: host doubles).  sb == 0 with bn == NULL never reads W.  If `dot_dev` is not NULL it receives
 * dot[j] = sum_i W_new[i,j]^2 in the same pass (e.g. the new squared residual norm of a CG step). */
int scimlop_col_axpby_dot(scimlop_handle_t h, int dtype, int64_t n, int64_t k, const void* X, int64_t ldx,
                          double sa, const void* an_dev, const void* ad_dev, void* W, int64_t ldw, double sb,
                          const void* bn_dev, const void* bd_dev, void* dot_dev, void* stream);

/* ---- multi-GPU: column sharding ---------------------------------------------------- */
/* Columns [first, first+count) of a k-column batch owned by `rank` of `nranks` (contiguous slabs). */
int scimlop_shard_columns(int64_t k, int nranks, int rank, int64_t* first, int64_t* count);
/* NCCL communicator over one process per GPU.  `unique_id` is the 128-byte ncclUniqueId produced by
 * scimlop_comm_unique_id on rank 0 and broadcast by the caller (torch.distributed / MPI / Julia Distributed). */
int scimlop_comm_unique_id(void* unique_id_128B);
int scimlop_comm_init(scimlop_handle_t h, int nranks, int rank, const void* unique_id_128B);
int scimlop_comm_destroy(scimlop_handle_t h);
/* In-place all-gather of equal column slabs: W_full is rows x (cols_per_rank*nranks) with ld == rows;
 * rank r's slab already sits at column r*cols_per_rank.  Only used when the caller wants a replicated W. */
int scimlop_allgather_cols(scimlop_handle_t h, int dtype, void* W_full, int64_t rows,
                           int64_t cols_per_rank, void* stream);

/* Column-sharded kron mul! whose result is wanted REPLICATED on every rank, with the all-gather overlapped
 * with the compute: this rank's `cols_per_rank` columns are processed in chunks of `chunk_cols` (0 = library
 * default); as soon as a chunk of W is done on the compute stream, the matching chunks of all ranks are
 * exchanged on a second stream (one grouped NCCL broadcast per rank over NVLink), while the next chunk computes.
 * V is this rank's slab (prod cols x cols_per_rank); W_full is prod rows x (cols_per_rank * nranks), contiguous,
 * rank r's columns at [r*cols_per_rank, (r+1)*cols_per_rank).  `workspace` must cover `chunk_cols` columns
 * (scimlop_kron_workspace_bytes with k = chunk_cols, or simply with k = cols_per_rank).  Returns after
 * enqueueing; `stream` is ordered after both the compute and the exchange. */
int scimlop_kron_mul_allgather(scimlop_handle_t h, int dtype, int precision, int nfactors,
                               const void* const* factors, const int64_t* dims, const int64_t* lds,
                               const int32_t* kinds, const void* V, void* W_full, int64_t cols_per_rank,
                               const void* alpha, int64_t chunk_cols, void* workspace, size_t workspace_bytes,
                               void* stream);

/* ---- fused compute + all-gather over peer memory (NVLink stores from the GEMM epilogue) ------------------------
 * One process per GPU; each rank exports its W_full allocation (scimlop_peer_export: a 64-byte CUDA IPC handle
 * plus the offset of the pointer inside its allocation), the handles travel through the caller's rendezvous,
 * and every rank maps the others' buffers (scimlop_peer_open).  scimlop_kron_mul_peers then computes this rank's
 * column slab ONCE and stores every output tile into its own W_full AND into the same columns of every peer's
 * W_full from the stage-2 epilogue, tile by tile while the tensor pipe works on the next tile -- no separate
 * collective.  scimlop_comm_barrier (a one-word NCCL all-reduce on the stream) tells every rank that all slabs
 * have landed.  Float64, outermost factor dense, beta == 0 (the result is overwritten).
 * ORDERING CONTRACT: the remote stores of call i+1 may only start once every rank has finished READING the W of call
 * i.  A caller that consumes W between calls must therefore also issue scimlop_comm_barrier BEFORE the next
 * scimlop_kron_mul_peers / _mc (after its reads of W are enqueued), or alternate between two W buffers. */
int scimlop_peer_export(scimlop_handle_t h, const void* dev_ptr, void* ipc_handle_64B, int64_t* offset);
int scimlop_peer_open(scimlop_handle_t h, const void* ipc_handle_64B, int64_t offset, void** peer_ptr);
int scimlop_peer_close(scimlop_handle_t h, void* peer_ptr, int64_t offset);
int scimlop_kron_mul_peers(scimlop_handle_t h, int dtype, int precision, int nfactors,
                           const void* const* factors, const int64_t* dims, const int64_t* lds,
                           const int32_t* kinds, const void* V, int64_t k, void* W_local,
                           void* const* peer_targets, int npeers, const void* alpha, void* workspace,
                           size_t workspace_bytes, void* stream);
int scimlop_comm_barrier(scimlop_handle_t h, void* stream);

/* ---- NVSwitch multicast windows (cuMulticast / NVLS) ------------------------------------------------------------
 * A buffer every rank of the node holds a copy of, plus a multicast alias through which ONE store lands in all
 * copies.  Collective set-up: scimlop_mc_open on every rank (rank 0 creates the object and returns a POSIX file
 * descriptor in *fd_out, which the caller hands to the other ranks -- SCM_RIGHTS over a Unix socket -- as their
 * fd_in), a barrier, scimlop_mc_bind on every rank (allocates and binds the local copy; returns the local array and
 * the multicast alias), a barrier.  scimlop_mc_padded_bytes rounds a size up to the multicast granularity. */
int scimlop_mc_padded_bytes(scimlop_handle_t h, size_t bytes, int nranks, size_t* padded);
int scimlop_mc_open(scimlop_handle_t h, size_t padded_bytes, int nranks, int rank, int fd_in, int* fd_out,
                    void** window);
int scimlop_mc_bind(scimlop_handle_t h, void* window, void** local_ptr, void** mc_ptr);
int scimlop_mc_free(scimlop_handle_t h, void* window);

/* Fused compute + all-gather through the switch: like scimlop_kron_mul_peers, but the last mode product stores every
 * output element once, with multimem.st, to `W_multicast` (the multicast alias of this rank's column slab); the
 * NVSwitch replicates it into every rank's copy (this one included), so the NVLink egress per GPU is one slab
 * instead of ranks-1 slabs.  Follow with scimlop_comm_barrier before any rank reads W. */
int scimlop_kron_mul_mc(scimlop_handle_t h, int dtype, int precision, int nfactors, const void* const* factors,
                        const int64_t* dims, const int64_t* lds, const int32_t* kinds, const void* V, int64_t k,
                        void* W_local, void* W_multicast, const void* alpha, void* workspace,
                        size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCIMLOP_B200_H */
