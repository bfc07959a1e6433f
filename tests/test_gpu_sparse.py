"""GPU parity of the sparse (CSR) MatrixOperator: `mul!(w, L.A, v[, α, β])` with a `sprand` matrix
(src/matrix.jl:337-350; test/basic.jl:405-426, test/Downstream/alloccheck.jl:37-45), through scimlop_csr_mul."""
import numpy as np
import pytest

from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu
TOL = {np.dtype("float64"): 1e-12, np.dtype("float32"): 1e-5}


@pytest.fixture(scope="module")
def sb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    import scimlop_b200 as sb
    sb.handle(0)
    return sb


def sprand(rng, m, n, density, dtype):
    A = rng.standard_normal((m, n)) * (rng.random((m, n)) < density)
    return np.asfortranarray(A.astype(dtype))


def csr_of(A):
    rows, cols = np.nonzero(A)
    rowptr = np.zeros(A.shape[0] + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=A.shape[0]), out=rowptr[1:])
    return rowptr, cols, A[rows, cols]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(100, 100, 5 / 100, 7), (100, 100, 2 / 100, 1), (257, 130, 0.05, 9), (64, 300, 0.5, 2),
                                   (1000, 1000, 0.01, 33), (5, 5, 0.0, 3)])
def test_sparse_matrix_operator_mul(sb, dtype, shape):
    m, n, density, k = shape
    rng = np.random.default_rng(m + n + k)
    A = sprand(rng, m, n, density, dtype)
    L = sb.MatrixOperator(sb.SparseMatrixCSR.from_dense(A))
    assert L.size() == (m, n)
    v = np.asfortranarray(rng.standard_normal((n, k)).astype(dtype))
    dv = sb.DeviceArray.from_numpy(v)
    w = sb.DeviceArray.full((m, k), float("nan"), dtype)
    sb.mul_(w, L, dv)
    got = w.numpy()
    assert np.isfinite(got).all(), "beta == 0 must not read W"
    rp, ci, va = csr_of(A)
    want = R.csr_mul(rp, ci, va, (m, n), v)
    assert R.rel_err(got, want) <= TOL[np.dtype(dtype)]
    assert R.rel_err(got, A.astype(np.float64) @ v.astype(np.float64)) <= TOL[np.dtype(dtype)]
    w0 = np.asfortranarray(rng.standard_normal((m, k)).astype(dtype))
    dw = sb.DeviceArray.from_numpy(w0)
    sb.mul_(dw, L, dv, 0.7, -1.5)
    assert R.rel_err(dw.numpy(), R.csr_mul(rp, ci, va, (m, n), v, 0.7, -1.5, w0)) <= TOL[np.dtype(dtype)]
    # lazy adjoint (src/matrix.jl:174-175): the scatter kernel
    Lt = L.adjoint()
    assert Lt.size() == (n, m)
    vt = np.asfortranarray(rng.standard_normal((m, k)).astype(dtype))
    dvt = sb.DeviceArray.from_numpy(vt)
    wt = sb.DeviceArray.full((n, k), float("nan"), dtype)
    sb.mul_(wt, Lt, dvt)
    assert np.isfinite(wt.numpy()).all()
    assert R.rel_err(wt.numpy(), R.csr_mul(rp, ci, va, (m, n), vt, trans=True)) <= TOL[np.dtype(dtype)]
    wt0 = np.asfortranarray(rng.standard_normal((n, k)).astype(dtype))
    dwt = sb.DeviceArray.from_numpy(wt0)
    sb.mul_(dwt, Lt, dvt, 2.0, 0.5)
    assert R.rel_err(dwt.numpy(), R.csr_mul(rp, ci, va, (m, n), vt, 2.0, 0.5, wt0, trans=True)) <= TOL[np.dtype(dtype)]


def test_sparse_operators_in_an_added_tree_with_time_dependent_scalars(sb):
    """test/basic.jl:405-426 / alloccheck.jl:37-60: MatrixOperator(A) + ScalarOperator(t) * MatrixOperator(A) + ..."""
    rng = np.random.default_rng(3)
    N, K = 100, 4
    A1, A2 = sprand(rng, N, N, 5 / N, np.float64), sprand(rng, N, N, 5 / N, np.float64)
    D = np.asfortranarray(rng.standard_normal((N, N)))
    c1 = sb.ScalarOperator(0.0, lambda a, u, p, t: np.sin(p * t))
    c2 = sb.ScalarOperator(0.0, lambda a, u, p, t: np.cos(p * t))
    op = sb.MatrixOperator(sb.SparseMatrixCSR.from_dense(A1)) + c1 * sb.MatrixOperator(sb.SparseMatrixCSR.from_dense(A2)) + \
        c2 * sb.MatrixOperator(D)
    v = np.asfortranarray(rng.standard_normal((N, K)))
    dv = sb.DeviceArray.from_numpy(v)
    op = sb.cache_operator(op, dv)
    for t in (0.3, 1.7):
        sb.update_coefficients_(op, None, 2.0, t)
        want = (A1 + np.sin(2.0 * t) * A2 + np.cos(2.0 * t) * D) @ v
        assert R.rel_err((op * dv).numpy(), want) <= 1e-12
        w0 = np.asfortranarray(rng.standard_normal((N, K)))
        dw = sb.DeviceArray.from_numpy(w0)
        sb.mul_(dw, op, dv, 0.5, 2.0)
        assert R.rel_err(dw.numpy(), 0.5 * want + 2.0 * w0) <= 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(300, 257, 40, 0.05), (300, 257, 128, 0.3), (1000, 500, 256, 0.004), (33, 2000, 384, 0.05)])
def test_cached_sparse_operator_uses_the_staged_gather(sb, dtype, shape):
    """cache_operator allocates the row-contiguous staging copy of V (k a multiple of 8): two launches, same numbers.
    k a multiple of 512 bytes of elements takes the warp-per-row kernel (csr_gather_rows_kernel): row counts that are not
    a multiple of the 32-row tile, rows with more than 32 nonzeros (density 0.3 of 257 columns), empty rows (0.004)."""
    rng = np.random.default_rng(8)
    m, n, k, density = shape
    A = sprand(rng, m, n, density, dtype)
    if density < 0.01:
        A[::7, :] = 0            # some rows without any nonzero
    v = np.asfortranarray(rng.standard_normal((n, k)).astype(dtype))
    dv = sb.DeviceArray.from_numpy(v)
    L = sb.cache_operator(sb.MatrixOperator(sb.SparseMatrixCSR.from_dense(A)), dv)
    assert L._csr_ws is not None
    h = sb.handle(0)
    w = sb.DeviceArray.full((m, k), float("nan"), dtype)
    sb.mul_(w, L, dv)
    n0 = h.launch_count()
    sb.mul_(w, L, dv)
    assert h.launch_count() - n0 == 2
    want = A.astype(np.float64) @ v.astype(np.float64)
    assert R.rel_err(w.numpy(), want) <= TOL[np.dtype(dtype)]
    w0 = np.asfortranarray(rng.standard_normal((m, k)).astype(dtype))
    dw = sb.DeviceArray.from_numpy(w0)
    sb.mul_(dw, L, dv, -0.5, 2.0)
    assert R.rel_err(dw.numpy(), -0.5 * want + 2.0 * w0.astype(np.float64)) <= TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.complex128])
@pytest.mark.parametrize("shape", [(300, 257, 40, 0.05), (257, 300, 128, 0.2), (64, 50, 3, 0.0)])
def test_cached_sparse_adjoint_gathers_on_the_transposed_structure(sb, dtype, shape):
    """cache_operator of the lazy adjoint (src/matrix.jl:174-175) builds the CSR form of A^T once; mul! is then the gather
    kernel on it: same numbers as the scatter kernel of the uncached adjoint within tolerance, and -- unlike the atomic
    scatter -- bit-identical from call to call.  Complex: the adjoint conjugates."""
    import torch
    rng = np.random.default_rng(21)
    m, n, k, density = shape
    A = sprand(rng, m, n, density, np.float64)
    if np.issubdtype(dtype, np.complexfloating):
        A = A + 1j * sprand(rng, m, n, density, np.float64) * (A != 0)
    A = np.asfortranarray(A.astype(dtype))
    real = np.float32 if dtype == np.float32 else np.float64
    v = rng.standard_normal((m, k)).astype(real)
    if np.issubdtype(dtype, np.complexfloating):
        v = v + 1j * rng.standard_normal((m, k))
    v = np.asfortranarray(v.astype(dtype))
    dv = sb.DeviceArray.from_numpy(v)
    L = sb.MatrixOperator(sb.SparseMatrixCSR.from_dense(A))
    plain = sb.DeviceArray.full((n, k), float("nan"), dtype)
    sb.mul_(plain, L.adjoint(), dv)                      # uncached: scatter kernel
    Lt = sb.cache_operator(L.adjoint(), dv)
    assert Lt._csr_T.shape == (n, m) and Lt._csr_T.nnz == L.A.nnz
    w = sb.DeviceArray.full((n, k), float("nan"), dtype)
    sb.mul_(w, Lt, dv)
    want = A.conj().T.astype(np.complex128 if np.issubdtype(dtype, np.complexfloating) else np.float64) @ v
    tol = 1e-5 if dtype == np.float32 else 1e-12

    def err(got, ref):          # norm-wise, on the complex values themselves (R.rel_err is for real arrays)
        ref = np.asarray(ref)
        return float(np.linalg.norm(np.asarray(got) - ref) / max(np.linalg.norm(ref), 1e-300)) if ref.size else 0.0
    assert err(w.numpy(), want) <= tol
    assert err(plain.numpy(), want) <= tol
    first = w.t.clone()
    for _ in range(3):
        w.t.fill_(float("nan"))
        sb.mul_(w, Lt, dv)
        assert torch.equal(w.t, first), "the gather on the transposed structure must be deterministic"
    if not np.issubdtype(dtype, np.complexfloating):
        w0 = np.asfortranarray(rng.standard_normal((n, k)).astype(dtype))
        dw = sb.DeviceArray.from_numpy(w0)
        sb.mul_(dw, Lt, dv, -2.0, 0.5)
        assert R.rel_err(dw.numpy(), -2.0 * want + 0.5 * w0.astype(np.float64)) <= tol


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", ["sparse_inner", "sparse_outer", "both", "three", "adjoint", "staged_inner"])
def test_sparse_kronecker_factors(sb, dtype, case):
    """`TensorProductOperator(sprand(...), B)`: a general sparse matrix as a Kronecker factor (SCIMLOP_CSR), along the
    innermost mode (the SpMM, staged when the workspace and the column count allow), along outer modes (gather over the
    contiguous mode rows, no copy), rectangular, next to dense / identity factors, as a lazy adjoint (converted once) --
    against the dense Kronecker definition (test/matrix.jl:625-644 style) through the oracle's mode products."""
    rng = np.random.default_rng(sum(map(ord, case)))
    tol = TOL[np.dtype(dtype)]
    S = lambda m, n, d: sprand(rng, m, n, d, dtype)
    D = lambda m, n: np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
    sp = lambda A: sb.MatrixOperator(sb.SparseMatrixCSR.from_dense(A))
    k = 5
    if case == "sparse_inner":
        mats = [D(7, 9), S(40, 33, 0.2)]; ops = [sb.MatrixOperator(mats[0]), sp(mats[1])]
    elif case == "sparse_outer":
        mats = [S(37, 41, 0.15), D(12, 10)]; ops = [sp(mats[0]), sb.MatrixOperator(mats[1])]
    elif case == "both":
        mats = [S(30, 30, 0.5), S(150, 140, 0.02)]; ops = [sp(mats[0]), sp(mats[1])]      # > 32 nonzeros in a row of the outer factor? 0.5 * 30 = 15; inner rows sparse incl. empty ones
        mats[0][3, :] = rng.standard_normal(30).astype(dtype); ops[0] = sp(mats[0])
    elif case == "three":
        mats = [S(6, 5, 0.6), np.asfortranarray(np.eye(4, dtype=dtype)), S(20, 25, 0.3)]
        ops = [sp(mats[0]), sb.IdentityOperator(4), sp(mats[2])]
    elif case == "adjoint":
        A0, B0 = S(21, 17, 0.3), S(19, 26, 0.2)
        mats = [np.asfortranarray(A0.T), np.asfortranarray(B0.T)]; ops = [sp(A0).adjoint(), sp(B0).adjoint()]
    else:   # inner factor with R = k * (columns of the outer factor) = 64: the staged, warp-per-row gather
        k = 8
        mats = [D(5, 8), S(300, 260, 0.1)]; ops = [sb.MatrixOperator(mats[0]), sp(mats[1])]
    ncols = int(np.prod([M.shape[1] for M in mats]))
    nrows = int(np.prod([M.shape[0] for M in mats]))
    v = np.asfortranarray(rng.standard_normal((ncols, k)).astype(dtype))
    dv = sb.DeviceArray.from_numpy(v)
    L = sb.cache_operator(sb.TensorProductOperator(*ops), dv)
    assert L._native is not None, "sparse factors must take the native Kronecker path"
    want = R.kron_apply([M.astype(np.float64) for M in mats], v.astype(np.float64))
    w = sb.DeviceArray.full((nrows, k), float("nan"), dtype)
    sb.mul_(w, L, dv)
    assert np.isfinite(w.numpy()).all()
    assert R.rel_err(w.numpy(), want) <= tol, f"{case}: {R.rel_err(w.numpy(), want):.3e}"
    w0 = np.asfortranarray(rng.standard_normal((nrows, k)).astype(dtype))
    dw = sb.DeviceArray.from_numpy(w0)
    sb.mul_(dw, L, dv, 0.5, -2.0)
    assert R.rel_err(dw.numpy(), 0.5 * want - 2.0 * w0.astype(np.float64)) <= tol
    # a vector input (k = 1) through a fresh cache
    v1 = np.asfortranarray(rng.standard_normal((ncols, 1)).astype(dtype))
    L1 = sb.cache_operator(sb.TensorProductOperator(*ops), sb.DeviceArray.from_numpy(v1))
    got1 = (L1 * sb.DeviceArray.from_numpy(v1)).numpy()
    assert R.rel_err(got1, R.kron_apply([M.astype(np.float64) for M in mats], v1.astype(np.float64))) <= tol


def test_sparse_from_scipy_and_large_batch(sb):
    sp = pytest.importorskip("scipy.sparse")
    rng = np.random.default_rng(5)
    n, k = 4096, 64
    A = sp.random(n, n, density=8 / n, format="csr", random_state=7, dtype=np.float64)
    L = sb.MatrixOperator(A)
    assert L.sparse and L.A.nnz == A.nnz
    v = np.asfortranarray(rng.standard_normal((n, k)))
    got = (L * sb.DeviceArray.from_numpy(v)).numpy()
    assert R.rel_err(got, A @ v) <= 1e-12
