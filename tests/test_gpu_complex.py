"""GPU parity for ComplexF32 / ComplexF64 operators (test/basic.jl:405-426 and test/Downstream/alloccheck.jl:37 iterate
`for T in (Float32, Float64, ComplexF32, ComplexF64)`): complex MatrixOperator, TensorProductOperator and Added / Scaled
trees of them through scimlop_kron_mul's complex planner (four planar real GEMMs per mode), against numpy."""
import numpy as np
import pytest

from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu
TOL = {np.dtype("complex128"): 1e-12, np.dtype("complex64"): 1e-5}


@pytest.fixture(scope="module")
def sb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    import scimlop_b200 as sb
    sb.handle(0)
    return sb


def crand(rng, *shape, dtype=np.complex128):
    return np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype))


def rel(x, y):
    return np.linalg.norm((x - y).ravel()) / max(np.linalg.norm(x.ravel()), np.linalg.norm(y.ravel()), 1e-300)


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("shape", [(100, 100, 7), (37, 53, 3), (130, 64, 1), (512, 300, 40)])
def test_complex_matrix_operator(sb, dtype, shape):
    m, n, k = shape
    rng = np.random.default_rng(m + n)
    A, v = crand(rng, m, n, dtype=dtype), crand(rng, n, k, dtype=dtype)
    L = sb.MatrixOperator(A)
    dv = sb.DeviceArray.from_numpy(v)
    w = sb.DeviceArray.from_numpy(np.full((m, k), np.nan + 1j * np.nan, dtype=dtype, order="F"))
    sb.mul_(w, L, dv)
    A64, v64 = A.astype(np.complex128), v.astype(np.complex128)
    assert np.isfinite(w.numpy()).all(), "beta == 0 must not read W"
    assert rel(w.numpy(), A64 @ v64) <= TOL[np.dtype(dtype)]
    w0 = crand(rng, m, k, dtype=dtype)
    dw = sb.DeviceArray.from_numpy(w0)
    alpha, beta = 0.7 - 0.2j, -1.3 + 0.5j
    sb.mul_(dw, L, dv, alpha, beta)
    assert rel(dw.numpy(), alpha * (A64 @ v64) + beta * w0.astype(np.complex128)) <= TOL[np.dtype(dtype)]
    # lazy adjoint = conjugate transpose (src/matrix.jl:174-175)
    vt = crand(rng, m, k, dtype=dtype)
    wt = (L.adjoint() * sb.DeviceArray.from_numpy(vt)).numpy()
    assert rel(wt, A64.conj().T @ vt.astype(np.complex128)) <= TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_complex_tensor_product(sb, dtype):
    rng = np.random.default_rng(5)
    A, B = crand(rng, 12, 9, dtype=dtype), crand(rng, 20, 16, dtype=dtype)
    k = 5
    v = crand(rng, 9 * 16, k, dtype=dtype)
    dv = sb.DeviceArray.from_numpy(v)
    L = sb.cache_operator(sb.TensorProductOperator(A, B), dv)
    want = np.kron(A.astype(np.complex128), B.astype(np.complex128)) @ v.astype(np.complex128)
    assert rel((L * dv).numpy(), want) <= TOL[np.dtype(dtype)]
    w0 = crand(rng, 12 * 20, k, dtype=dtype)
    dw = sb.DeviceArray.from_numpy(w0)
    sb.mul_(dw, L, dv, 2.0 + 1.0j, 0.5j)
    assert rel(dw.numpy(), (2.0 + 1.0j) * want + 0.5j * w0.astype(np.complex128)) <= TOL[np.dtype(dtype)]
    # three factors with an identity in the middle and an adjoint factor
    Cm = crand(rng, 7, 7, dtype=dtype)
    v3 = crand(rng, 9 * 4 * 7, 3, dtype=dtype)
    dv3 = sb.DeviceArray.from_numpy(v3)
    L3 = sb.cache_operator(sb.TensorProductOperator(sb.MatrixOperator(A), sb.IdentityOperator(4), sb.MatrixOperator(Cm).adjoint()), dv3)
    want3 = np.kron(np.kron(A.astype(np.complex128), np.eye(4)), Cm.astype(np.complex128).conj().T) @ v3.astype(np.complex128)
    assert rel((L3 * dv3).numpy(), want3) <= TOL[np.dtype(dtype)]


def test_complex_added_tree_with_time_dependent_scalars(sb):
    """test/basic.jl:405-426: Op = -1im * (O1 - O2) with O_i = MatrixOperator(A) + ScalarOperator(t^p) * MatrixOperator(A)."""
    rng = np.random.default_rng(6)
    N, K = 100, 4
    A = crand(rng, N, N)
    s1 = sb.ScalarOperator(0.0, lambda a, u, p, t: t)
    s2 = sb.ScalarOperator(0.0, lambda a, u, p, t: t ** 2)
    O1 = sb.MatrixOperator(A) + s1 * sb.MatrixOperator(A)
    O2 = sb.MatrixOperator(A) + s2 * sb.MatrixOperator(A)
    Op = -1j * (O1 + (-1.0) * O2)
    v = crand(rng, N, K)
    dv = sb.DeviceArray.from_numpy(v)
    Op = sb.cache_operator(Op, dv)
    for t in (0.5, 2.0):
        sb.update_coefficients_(Op, None, None, t)
        want = -1j * ((1 + t) * A - (1 + t ** 2) * A) @ v
        assert rel((Op * dv).numpy(), want) <= 1e-12
        w0 = crand(rng, N, K)
        dw = sb.DeviceArray.from_numpy(w0)
        sb.mul_(dw, Op, dv, 0.5 + 0.5j, 2.0)
        assert rel(dw.numpy(), (0.5 + 0.5j) * want + 2.0 * w0) <= 1e-12


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_complex_sparse_matrix_operator(sb, dtype):
    """alloccheck.jl:37-45: MatrixOperator(sprand(T, N, N, 5 / N)) for complex T, next to a dense one in an AddedOperator."""
    rng = np.random.default_rng(9)
    N, K = 100, 6
    A = crand(rng, N, N, dtype=dtype) * (rng.random((N, N)) < 5 / N)
    A = np.asfortranarray(A.astype(dtype))
    D = crand(rng, N, N, dtype=dtype)
    v = crand(rng, N, K, dtype=dtype)
    dv = sb.DeviceArray.from_numpy(v)
    Ls = sb.MatrixOperator(sb.SparseMatrixCSR.from_dense(A))
    w = sb.DeviceArray.from_numpy(np.full((N, K), np.nan + 1j * np.nan, dtype=dtype, order="F"))
    sb.mul_(w, Ls, dv)
    A64, v64 = A.astype(np.complex128), v.astype(np.complex128)
    assert np.isfinite(w.numpy()).all()
    assert rel(w.numpy(), A64 @ v64) <= TOL[np.dtype(dtype)]
    w0 = crand(rng, N, K, dtype=dtype)
    dw = sb.DeviceArray.from_numpy(w0)
    sb.mul_(dw, Ls, dv, 0.5 - 1.0j, 2.0j)
    assert rel(dw.numpy(), (0.5 - 1.0j) * (A64 @ v64) + 2.0j * w0.astype(np.complex128)) <= TOL[np.dtype(dtype)]
    # transpose flag of the scatter kernel = plain transpose of the stored values; the lazy ADJOINT of a complex sparse
    # operator conjugates as well, which the host mirror does by conjugating a copy of the values
    vt = crand(rng, N, K, dtype=dtype)
    wt = (Ls.adjoint() * sb.DeviceArray.from_numpy(vt)).numpy()
    assert rel(wt, A64.conj().T @ vt.astype(np.complex128)) <= TOL[np.dtype(dtype)]
    op = sb.cache_operator(Ls + (1.0 + 0.5j) * sb.MatrixOperator(D), dv)
    assert rel((op * dv).numpy(), (A64 + (1.0 + 0.5j) * D.astype(np.complex128)) @ v64) <= TOL[np.dtype(dtype)]
