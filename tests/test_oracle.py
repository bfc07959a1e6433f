"""Pins the CPU oracle (oracle/) against the reference's own test definitions:
the dense materialisation `kron(A,B[,C]) * v`, the exact known answers of
test/precompile_workload.jl, the shape matrix of test/matrix.jl:510-517, the identity placements of
test/matrix.jl:696-703 and the tree cases of test/basic.jl.  CPU only."""
import numpy as np
import pytest

from oracle import ref_numpy as R

RTOL = 1e-12


def rnd(rng, *shape, dtype=np.float64):
    return np.asfortranarray(rng.random(shape).astype(dtype))


# ---- exact known answers: test/precompile_workload.jl:10-37 -------------------------------------
def test_known_answers_exact():
    v = np.array([1.0, 2.0])
    A = R.MatrixOperator(np.array([[2.0, 1.0], [1.0, 3.0]]))
    D = R.DiagonalOperator(np.array([2.0, 3.0]))
    w = np.zeros(2)
    assert (A.mul3(w, v) == [4.0, 7.0]).all()                                  # :10-14
    w = np.zeros(2)
    assert (R.AddedOperator(A, D).mul3(w, v) == [6.0, 13.0]).all()             # :22
    AD = R.ComposedOperator(A, D).cache_operator(v)
    w = np.zeros(2)
    assert (AD.mul3(w, v) == [10.0, 20.0]).all()                               # :23,:33-37
    w = np.zeros(2)
    assert (R.ScaledOperator(2.0, A).mul3(w, v) == [8.0, 14.0]).all()          # :25
    T = R.TensorProductOperator(R.IdentityOperator(2), D)
    v4 = np.array([1.0, 2.0, 3.0, 4.0])
    T = T.cache_operator(v4)
    w = np.zeros(4)
    assert (T.mul3(w, v4) == [2.0, 6.0, 6.0, 12.0]).all()                      # :30-31


# ---- TPO shape matrix: test/matrix.jl:510-517, K=19 (:12), 3-arg :625-631, 5-arg :637-644 --------
@pytest.mark.parametrize("square", [False, True])
@pytest.mark.parametrize("lv", [False, True])
@pytest.mark.parametrize("vector_input", [False, True])
def test_tpo_matches_dense_kron(square, lv, vector_input):
    rng = np.random.default_rng(0)
    m1, n1, m2, n2, m3, n3 = 3, 5, 7, 11, 13, 17
    if square:
        n1, n2, n3 = m1, m2, m3
    K = 19
    A, B, C = rnd(rng, m1, n1), rnd(rng, m2, n2), rnd(rng, m3, n3)
    alpha, beta = rng.random(), rng.random()
    for mats in ((A, B), (A, B, C)):
        dense = mats[0]
        for M in mats[1:]:
            dense = np.kron(dense, M)
        n, m = dense.shape[1], dense.shape[0]
        v = rnd(rng, n) if vector_input else rnd(rng, n, K)
        op = R.TensorProductOperator(*mats, lv=lv).cache_operator(v)
        assert np.allclose(op.to_dense(), dense, rtol=RTOL, atol=0)
        w = np.zeros((m,) + v.shape[1:], order="F")
        op.mul3(w, v)
        assert R.rel_err(w, dense @ v) < RTOL
        w0 = rnd(rng, *w.shape)
        w = w0.copy(order="F")
        op.mul5(w, v, alpha, beta)
        assert R.rel_err(w, alpha * (dense @ v) + beta * w0) < RTOL
        # β == 0 must not read w (src/batch.jl:222-223, BLAS convention at src/matrix.jl:349)
        w = np.full_like(w0, np.nan)
        op.mul5(w, v, alpha, 0.0)
        assert R.rel_err(w, alpha * (dense @ v)) < RTOL
        # structured oracle used at bench sizes
        assert R.rel_err(R.kron_apply(list(mats), v), dense @ v) < RTOL


# ---- identity placements: test/matrix.jl:696-703,755-758 ----------------------------------------
def test_tpo_identity_placements():
    rng = np.random.default_rng(1)
    m1, m2, m3, K = 3, 7, 13, 5
    n1, n2 = 3, 7
    A1, A2 = rnd(rng, n1, n1), rnd(rng, n2, n2)
    I = R.IdentityOperator
    cases = [
        ([A1, I(m1), I(m2), I(m3)]), ([I(m1), A1, I(m2), I(m3)]), ([I(m1), I(m2), A1, I(m3)]),
        ([I(m1), I(m2), I(m3), A1]), ([A1, A2, I(m1), I(m2)]), ([I(m1), A1, A2, I(m2)]),
        ([I(m1), I(m2), A1, A2]),
    ]
    for ops in cases:
        dense = np.eye(1)
        for o in ops:
            dense = np.kron(dense, o if isinstance(o, np.ndarray) else np.eye(o.len))
        for lv in (False, True):
            op = R.TensorProductOperator(*ops, lv=lv)
            v = rnd(rng, dense.shape[1], K)
            op = op.cache_operator(v)
            w = np.zeros((dense.shape[0], K), order="F")
            op.mul3(w, v)
            assert R.rel_err(w, dense @ v) < RTOL


# ---- TensorSum: test/matrix.jl:809-867 ------------------------------------------------------------
def test_tensor_sum():
    rng = np.random.default_rng(2)
    A, B = rnd(rng, 5, 5), rnd(rng, 7, 7)
    dense = np.kron(A, np.eye(7)) + np.kron(np.eye(5), B)
    for K in (1, 6):
        v = rnd(rng, 35, K)
        op = R.TensorSumOperator(A, B).cache_operator(v)
        w = np.zeros((35, K), order="F")
        op.mul3(w, v)
        assert R.rel_err(w, dense @ v) < RTOL
        w0 = rnd(rng, 35, K)
        w = w0.copy(order="F")
        op.mul5(w, v, 0.7, 0.3)
        assert R.rel_err(w, 0.7 * dense @ v + 0.3 * w0) < RTOL


# ---- trees: test/basic.jl:258-318 (Scaled), :320-384 (Added), :516-608 (Composed) ----------------
def test_added_scaled_composed():
    rng = np.random.default_rng(3)
    N, K = 8, 12
    A, B, C = (R.MatrixOperator(rnd(rng, N, N)) for _ in range(3))
    a, b = rng.random(), rng.random()
    v = rnd(rng, N, K)
    for sgn in (1.0, -1.0):
        for op, dense in (
            (R.AddedOperator(A, R.ScaledOperator(sgn, B)), A.A + sgn * B.A),
            (R.AddedOperator(R.ScaledOperator(a, A), R.ScaledOperator(sgn * b, B)), a * A.A + sgn * b * B.A),
        ):
            w = np.zeros((N, K), order="F")
            op.mul3(w, v)
            assert R.rel_err(w, dense @ v) < RTOL
            w0 = rnd(rng, N, K)
            w = w0.copy(order="F")
            op.mul5(w, v, a, b)
            assert R.rel_err(w, a * dense @ v + b * w0) < RTOL
    comp = R.ComposedOperator(A, B, C).cache_operator(v)
    dense = A.A @ B.A @ C.A
    w = np.zeros((N, K), order="F")
    comp.mul3(w, v)
    assert R.rel_err(w, dense @ v) < RTOL
    w0 = rnd(rng, N, K)
    w = w0.copy(order="F")
    comp.mul5(w, v, a, b)
    assert R.rel_err(w, a * dense @ v + b * w0) < RTOL


# ---- config-4 tree α(D*M' + 2M + I)v + βw with Diagonal / BatchedDiagonal: README.md:69-73 style --
def test_config4_tree_and_batched_diag():
    rng = np.random.default_rng(4)
    N, K = 16, 7
    M = R.MatrixOperator(rnd(rng, N, N))
    D = R.DiagonalOperator(rnd(rng, N))
    Bd = R.BatchedDiagonalOperator(rnd(rng, N, K))
    v = rnd(rng, N, K)
    L = R.AddedOperator(R.ComposedOperator(D, M.adjoint()), R.ScaledOperator(2.0, M),
                        R.ScaledOperator(True, R.IdentityOperator(N))).cache_operator(v)
    dense = np.diag(D.d) @ M.A.T + 2 * M.A + np.eye(N)
    w0 = rnd(rng, N, K)
    w = w0.copy(order="F")
    L.mul5(w, v, 0.7, 0.3)
    assert R.rel_err(w, 0.7 * dense @ v + 0.3 * w0) < RTOL
    # elementwise tree with a batched diagonal (C4-elt)
    L2 = R.AddedOperator(R.ComposedOperator(D, Bd), R.ScaledOperator(2.0, Bd), R.IdentityOperator(N))
    L2 = L2.cache_operator(v)
    expect = (D.d[:, None] * Bd.diag + 2 * Bd.diag + 1.0) * v
    w = np.full((N, K), np.nan, order="F")
    L2.mul5(w, v, 0.7, 0.0)
    assert R.rel_err(w, 0.7 * expect) < RTOL
    assert R.rel_err(R.dense_apply(L2, v), expect) < RTOL
    w = w0.copy(order="F")
    L2.mul5(w, v, 0.7, 0.3)
    assert R.rel_err(w, 0.7 * expect + 0.3 * w0) < RTOL


# ---- C restatement agrees with the NumPy restatement and the definition -------------------------
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_c_oracle(oracle_c, dtype, tol):
    rng = np.random.default_rng(5)
    for (mo, no, mi, ni, k) in ((3, 5, 7, 11, 19), (7, 7, 3, 3, 1), (12, 12, 12, 12, 100), (5, 4, 70, 9, 3)):
        A, B = rnd(rng, mo, no, dtype=dtype), rnd(rng, mi, ni, dtype=dtype)
        v = rnd(rng, no * ni, k, dtype=dtype)
        w0 = rnd(rng, mo * mi, k, dtype=dtype)
        dense = np.kron(A.astype(np.float64), B.astype(np.float64))
        for lv in (True, False):
            w = np.zeros_like(w0)
            oracle_c.tpo2_mul(w, A, B, v, lv=lv)
            assert R.rel_err(w, dense @ v) < tol
            w = w0.copy(order="F")
            oracle_c.tpo2_mul(w, A, B, v, 0.7, 0.3, lv=lv)
            assert R.rel_err(w, 0.7 * dense @ v + 0.3 * w0) < tol
            w = np.full_like(w0, np.nan)
            oracle_c.tpo2_mul(w, A, B, v, 0.7, 0.0, lv=lv)
            assert R.rel_err(w, 0.7 * dense @ v) < tol
    d = rnd(rng, 33, dtype=dtype)
    bd = rnd(rng, 33, 5, dtype=dtype)
    v = rnd(rng, 33, 5, dtype=dtype)
    w0 = rnd(rng, 33, 5, dtype=dtype)
    w = w0.copy(order="F")
    oracle_c.diag_mul(w, d, v, 0.7, 0.3)
    assert R.rel_err(w, 0.7 * d[:, None] * v + 0.3 * w0) < tol
    w = np.full_like(w0, np.nan)
    oracle_c.diag_mul(w, bd, v, 0.7, 0.0)
    assert R.rel_err(w, 0.7 * bd * v) < tol


@pytest.mark.parametrize("vector_input", [False, True])
def test_tpo_ldiv_restatement_matches_dense_solve(vector_input):
    """test/matrix.jl:646-671, :801-806: `ldiv!(w, factorize(opAB), v) ≈ AB \\ v` for the square shapes."""
    rng = np.random.default_rng(5)
    for n1, n2, n3 in [(3, 5, 7), (7, 4, 2)]:
        A, B, Cm = (np.asfortranarray(rng.random((n, n)) + np.eye(n)) for n in (n1, n2, n3))
        K = 1 if vector_input else 19
        v = rng.random(n1 * n2) if vector_input else np.asfortranarray(rng.random((n1 * n2, K)))
        want = np.linalg.solve(np.kron(A, B), v)
        assert R.rel_err(R.tpo_ldiv(A, B, v), want) < 1e-12
        assert R.rel_err(R.kron_solve([A, B], v), want) < 1e-12
        v3 = rng.random(n1 * n2 * n3) if vector_input else np.asfortranarray(rng.random((n1 * n2 * n3, K)))
        want3 = np.linalg.solve(np.kron(A, np.kron(B, Cm)), v3)
        assert R.rel_err(R.kron_solve([A, B, Cm], v3), want3) < 1e-11
        # nested like the reference's A ⊗ (B ⊗ C): the inner solve is itself a TPO ldiv!
        assert R.rel_err(R.tpo_ldiv(A, np.kron(B, Cm), v3), want3) < 1e-11


def test_block_diagonal_and_woperator_restatements():
    """src/block.jl:145-179 and src/woperator.jl:447-481 against their dense definitions
    (test/block.jl: `L * v ≈ Matrix(L) * v`; test/woperator.jl: `W * v ≈ (J - M/γ) * v`)."""
    import scipy.linalg as sla
    rng = np.random.default_rng(9)
    blocks = [rnd(rng, 3, 5), np.diag(rng.random(4)), rnd(rng, 6, 2), np.eye(3)]
    dense = sla.block_diag(*blocks)
    v = rnd(rng, dense.shape[1], 7)
    w0 = rnd(rng, dense.shape[0], 7)
    assert R.rel_err(R.block_diag_mul(np.full_like(w0, np.nan), blocks, v), dense @ v) < RTOL
    assert R.rel_err(R.block_diag_mul(w0.copy(), blocks, v, 0.7, 0.3), 0.7 * dense @ v + 0.3 * w0) < RTOL
    n = 9
    J, M, gamma = rnd(rng, n, n), rnd(rng, n, n), 0.37
    b = rng.random(n)
    assert R.rel_err(R.woperator_mul(np.empty(n), M, gamma, J, b), (J - M / gamma) @ b) < RTOL
    assert R.rel_err(R.woperator_mul(np.empty(n), 2.0, gamma, J, b), (J - 2.0 * np.eye(n) / gamma) @ b) < RTOL


def test_csr_restatement_matches_the_dense_product():
    """The reference's own check for sparse MatrixOperators: L * v ≈ Matrix(A) * v (test/basic.jl:405-426)."""
    rng = np.random.default_rng(21)
    for m, n, k in ((7, 5, 3), (40, 40, 1), (1, 9, 2), (12, 30, 4)):
        A = rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.3)
        rows, cols = np.nonzero(A)
        rowptr = np.zeros(m + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=m), out=rowptr[1:])
        vals = A[rows, cols]
        v = rng.standard_normal((n, k))
        assert np.allclose(R.csr_mul(rowptr, cols, vals, (m, n), v), A @ v, rtol=1e-13, atol=1e-13)
        w0 = rng.standard_normal((m, k))
        assert np.allclose(R.csr_mul(rowptr, cols, vals, (m, n), v, 0.3, -2.0, w0), 0.3 * (A @ v) - 2.0 * w0, rtol=1e-13, atol=1e-13)
        vt = rng.standard_normal((m, k))
        assert np.allclose(R.csr_mul(rowptr, cols, vals, (m, n), vt, trans=True), A.T @ vt, rtol=1e-13, atol=1e-13)
