"""GPU tests of the factorised solve (csrc/lu.cu): `factorize` -> blocked LU on the device, `ldiv!` as substitution
sweeps along the Kronecker modes.  Reference: `lu` / `ldiv!` of src/matrix.jl:513-516, src/tensor.jl:227, :612-673,
:864-899; test/matrix.jl:646-687.  A solve's FORWARD error scales with cond(A) for LAPACK's getrs too, so parity at high
condition numbers is stated as LAPACK states it: the normwise BACKWARD error  |b - A x| / (|A| |x| + |b|)  must be a small
multiple of eps at any cond -- which multiplication by an explicit inverse does not deliver."""
import numpy as np
import pytest

from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def sb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    import scimlop_b200 as sb
    sb.handle(0)
    return sb


def with_cond(rng, n, kappa):
    """Random matrix with 2-norm condition number kappa (geometric singular values)."""
    U, _ = np.linalg.qr(rng.standard_normal((n, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    s = np.geomspace(1.0, 1.0 / kappa, n)
    return np.asfortranarray((U * s) @ V.T)


def backward_error(A, x, b):
    r = b - A @ x
    return np.linalg.norm(r, np.inf) / (np.linalg.norm(A, np.inf) * np.linalg.norm(x, np.inf) + np.linalg.norm(b, np.inf))


@pytest.mark.parametrize("n", [1, 2, 5, 31, 32, 33, 100, 128, 129, 300, 512, 1000])
def test_factorize_inverse_and_condition_number(sb, n):
    rng = np.random.default_rng(n)
    A = np.asfortranarray(rng.standard_normal((n, n)))
    if n > 1:
        A[0, 0] = 0.0                       # the first pivot has to be exchanged
    lu = sb.LUFactorization(sb.DeviceArray.from_numpy(A))
    Ainv = lu.inv.numpy()
    kappa1 = np.linalg.norm(A, 1) * np.linalg.norm(np.linalg.inv(A), 1)
    assert abs(lu.cond - kappa1) <= 1e-6 * kappa1, f"cond_1 {lu.cond:.6e} vs {kappa1:.6e}"
    resid = np.abs(A @ Ainv - np.eye(n)).max()
    assert resid <= 50 * kappa1 * EPS, f"|A inv(A) - I| = {resid:.3e}, cond {kappa1:.2e}"


def test_factorize_singular_raises(sb):
    with pytest.raises(sb.ScimlopError) as e:
        sb.LUFactorization(sb.DeviceArray.from_numpy(np.asfortranarray(np.ones((40, 40)))))
    assert e.value.status == sb._lib.ERR_SINGULAR


@pytest.mark.parametrize("kappa", [1e2, 1e6, 1e10, 1e13])
@pytest.mark.parametrize("n,k", [(96, 7), (512, 64), (700, 33)])
def test_ill_conditioned_solve_is_backward_stable(sb, kappa, n, k):
    """test/matrix.jl:646-687 at cond = 1e2 ... 1e13: backward error of ldiv! within a small multiple of eps, like
    numpy.linalg.solve (LAPACK gesv) on the same system."""
    rng = np.random.default_rng(int(np.log10(kappa)) + n)
    A = with_cond(rng, n, kappa)
    b = np.asfortranarray(rng.standard_normal((n, k)))
    b[:, : k // 2] = A @ rng.standard_normal((n, k // 2))     # half of the columns: solutions of size ~1 (the hard case)
    L = sb.factorize(sb.MatrixOperator(A))
    assert L.stable_solve == (L.lu.cond > 1e3)
    db = sb.DeviceArray.from_numpy(b)
    x = sb.DeviceArray.full((n, k), float("nan"), np.float64)
    sb.ldiv_(x, L, db)
    got = x.numpy()
    assert np.isfinite(got).all()
    xl = np.linalg.solve(A, b)
    for j in range(k):      # column by column: a huge column must not hide a small one
        be = backward_error(A, got[:, j:j + 1], b[:, j:j + 1])
        be_lapack = backward_error(A, xl[:, j:j + 1], b[:, j:j + 1])
        assert be <= max(50 * EPS, 20 * be_lapack), f"cond {kappa:.0e} column {j}: backward error {be:.3e} (LAPACK {be_lapack:.3e})"
    be_lapack = backward_error(A, xl, b)
    # forward error against LAPACK's solution: both are within cond * eps of the truth
    assert R.rel_err(got, np.linalg.solve(A, b)) <= 200 * kappa * EPS
    # two-argument form
    sb.ldiv_(L, db)
    assert backward_error(A, db.numpy(), b) <= max(50 * EPS, 20 * be_lapack)


def test_explicit_inverse_would_fail_the_same_bound(sb):
    """Why the factorised path exists: for a right-hand side b = A x0 with |x0| ~ 1 (the solution does not live in the
    small singular directions) the inverse-multiply residual is ~cond * eps at cond = 1e10."""
    rng = np.random.default_rng(0)
    n, k = 256, 8
    A = with_cond(rng, n, 1e10)
    b = np.asfortranarray(A @ rng.standard_normal((n, k)))
    lu = sb.LUFactorization(sb.DeviceArray.from_numpy(A))
    x_inv = lu.inv.numpy() @ b
    L = sb.factorize(sb.MatrixOperator(A))
    x = sb.DeviceArray.full((n, k), float("nan"), np.float64)
    sb.ldiv_(x, L, sb.DeviceArray.from_numpy(b))
    assert backward_error(A, x.numpy(), b) <= 50 * EPS
    assert backward_error(A, x_inv, b) >= 100 * backward_error(A, x.numpy(), b)


@pytest.mark.parametrize("shape", [((64, 1e8), (96, 10.0), 5), ((130, 5.0), (128, 1e9), 3), ((512, 1e7), (512, 1e7), 4)])
def test_kron_ldiv_with_ill_conditioned_factors(sb, shape):
    """src/tensor.jl:612-673 / outer_div! :864-899 with the substitution sweeps acting along the inner and the outer mode:
    (A (x) B) \\ v against the mode-wise LAPACK solves of the oracle, backward error of the Kronecker system."""
    (na, ka), (nb, kb), k = shape
    rng = np.random.default_rng(na + nb)
    A, B = with_cond(rng, na, ka), with_cond(rng, nb, kb)
    v = np.asfortranarray(rng.standard_normal((na * nb, k)))
    dv = sb.DeviceArray.from_numpy(v)
    L = sb.cache_operator(sb.factorize(sb.TensorProductOperator(A, B)), dv)
    x = sb.DeviceArray.full((na * nb, k), float("nan"), np.float64)
    sb.ldiv_(x, L, dv)
    got = x.numpy()
    assert np.isfinite(got).all()
    # residual of the Kronecker system through the oracle's mul!
    r = v - R.kron_apply([A, B], got)
    scale = np.linalg.norm(A, np.inf) * np.linalg.norm(B, np.inf) * np.abs(got).max() + np.abs(v).max()
    assert np.abs(r).max() / scale <= 200 * EPS, f"Kronecker backward error {np.abs(r).max() / scale:.3e}"
    # mode-wise LAPACK solves (the reference's algorithm): same solution within cond * eps
    X = got.reshape(nb, na, k, order="F")
    want = np.empty_like(X)
    Vt = v.reshape(nb, na, k, order="F")
    for j in range(k):
        want[:, :, j] = np.linalg.solve(A, np.linalg.solve(B, Vt[:, :, j]).T).T
    assert R.rel_err(X, want) <= 500 * ka * kb * EPS
    # two-argument (in-place) form
    dv2 = sb.DeviceArray.from_numpy(v)
    sb.ldiv_(L, dv2)
    assert np.array_equal(dv2.numpy(), got)


@pytest.mark.parametrize("kappa", [1e6, 1e10])
@pytest.mark.parametrize("n,k", [(96, 7), (300, 5), (512, 64), (700, 33)])
def test_adjoint_of_ill_conditioned_factorisation_solves_with_transposed_sweeps(sb, kappa, n, k):
    """`adjoint(::InvertibleOperator)` (src/matrix.jl:565) shares the factorisation; its ldiv! is LAPACK getrs('T'): U^T
    and L^T sweeps, then P^T (scimlop_kron_ldiv with SCIMLOP_DENSE | SCIMLOP_TRANS).  Backward error of A^T x = b like
    numpy.linalg.solve(A.T, b)."""
    rng = np.random.default_rng(int(np.log10(kappa)) * 7 + n)
    A = with_cond(rng, n, kappa)
    At = np.ascontiguousarray(A.T)
    b = np.asfortranarray(rng.standard_normal((n, k)))
    b[:, : k // 2] = At @ rng.standard_normal((n, k // 2))
    L = sb.factorize(sb.MatrixOperator(A))
    Lt = L.adjoint()
    assert Lt.stable_solve and Lt.lu is L.lu
    db = sb.DeviceArray.from_numpy(b)
    x = sb.DeviceArray.full((n, k), float("nan"), np.float64)
    sb.ldiv_(x, Lt, db)
    got = x.numpy()
    assert np.isfinite(got).all()
    xl = np.linalg.solve(At, b)
    for j in range(k):
        be = backward_error(At, got[:, j:j + 1], b[:, j:j + 1])
        be_lapack = backward_error(At, xl[:, j:j + 1], b[:, j:j + 1])
        assert be <= max(50 * EPS, 20 * be_lapack), f"cond {kappa:.0e} column {j}: backward error {be:.3e} (LAPACK {be_lapack:.3e})"
    assert R.rel_err(got, xl) <= 200 * kappa * EPS
    # mul! of the adjoint is A^T v; the forward operator still solves with A
    w = sb.DeviceArray.full((n, k), float("nan"), np.float64)
    sb.mul_(w, Lt, db)
    assert R.rel_err(w.numpy(), At @ b) <= 1e-12
    sb.ldiv_(x, L, db)
    assert backward_error(A, x.numpy(), b) <= 1e3 * EPS
    # two-argument form
    sb.ldiv_(Lt, db)
    assert np.array_equal(db.numpy(), got)


@pytest.mark.parametrize("which", ["both", "outer", "inner"])
def test_kron_ldiv_with_transposed_factors(sb, which):
    """(A (x) B)' \\ v = (A' (x) B') \\ v (src/tensor.jl:181-191 + :612-673) and the mixed cases, ill-conditioned factors:
    transposed sweeps along the inner (first applied: V is copied, the sweeps consume their source) and the outer mode."""
    rng = np.random.default_rng(len(which))
    na, nb, k = 130, 96, 5
    A, B = with_cond(rng, na, 1e8), with_cond(rng, nb, 1e6)
    v = np.asfortranarray(rng.standard_normal((na * nb, k)))
    dv = sb.DeviceArray.from_numpy(v)
    FA, FB = sb.factorize(sb.MatrixOperator(A)), sb.factorize(sb.MatrixOperator(B))
    if which == "both":
        L = sb.factorize(sb.TensorProductOperator(A, B)).adjoint()
        Ae, Be = A.T, B.T
    elif which == "outer":
        L = sb.TensorProductOperator(FA.adjoint(), FB)
        Ae, Be = A.T, B
    else:
        L = sb.TensorProductOperator(FA, FB.adjoint())
        Ae, Be = A, B.T
    Ae, Be = np.asfortranarray(Ae), np.asfortranarray(Be)
    L = sb.cache_operator(L, dv)
    x = sb.DeviceArray.full((na * nb, k), float("nan"), np.float64)
    sb.ldiv_(x, L, dv)
    got = x.numpy()
    assert np.isfinite(got).all()
    r = v - R.kron_apply([Ae, Be], got)
    scale = np.linalg.norm(Ae, np.inf) * np.linalg.norm(Be, np.inf) * np.abs(got).max() + np.abs(v).max()
    assert np.abs(r).max() / scale <= 200 * EPS, f"{which}: Kronecker backward error {np.abs(r).max() / scale:.3e}"
    X = got.reshape(nb, na, k, order="F")
    Vt = v.reshape(nb, na, k, order="F")
    want = np.empty_like(X)
    for j in range(k):
        want[:, :, j] = np.linalg.solve(Ae, np.linalg.solve(Be, Vt[:, :, j]).T).T
    assert R.rel_err(X, want) <= 500 * 1e8 * 1e6 * EPS
    assert np.array_equal(dv.numpy(), v), "ldiv! modified its right-hand side"
    # the operator itself applies the transposed factors
    w = sb.DeviceArray.full((na * nb, k), float("nan"), np.float64)
    sb.mul_(w, L, dv)
    assert R.rel_err(w.numpy(), R.kron_apply([Ae, Be], v)) <= 1e-12
    dv2 = sb.DeviceArray.from_numpy(v)
    sb.ldiv_(L, dv2)
    assert np.array_equal(dv2.numpy(), got)


def test_three_factor_ldiv_with_identity_and_scaling(sb):
    rng = np.random.default_rng(9)
    A, Cm = with_cond(rng, 40, 1e8), with_cond(rng, 48, 1e6)
    k = 3
    v = np.asfortranarray(rng.standard_normal((40 * 16 * 48, k)))
    dv = sb.DeviceArray.from_numpy(v)
    L = sb.cache_operator(sb.factorize(sb.TensorProductOperator(2.0 * sb.MatrixOperator(A), sb.IdentityOperator(16), sb.MatrixOperator(Cm))), dv)
    x = sb.DeviceArray.full(v.shape, float("nan"), np.float64)
    sb.ldiv_(x, L, dv)
    r = v - 2.0 * R.kron_apply([A, np.eye(16), Cm], x.numpy())
    scale = 2.0 * np.linalg.norm(A, np.inf) * np.linalg.norm(Cm, np.inf) * np.abs(x.numpy()).max() + np.abs(v).max()
    assert np.abs(r).max() / scale <= 200 * EPS


def test_well_conditioned_factors_keep_the_inverse_fast_path(sb):
    rng = np.random.default_rng(3)
    n, k = 128, 6
    A = np.asfortranarray(rng.random((n, n)) + n * np.eye(n))
    L = sb.factorize(sb.MatrixOperator(A))
    assert L.lu.cond < 1e3 and not L.stable_solve
    b = np.asfortranarray(rng.standard_normal((n, k)))
    h = sb.handle(0)
    x = sb.DeviceArray.full((n, k), float("nan"), np.float64)
    sb.ldiv_(x, L, sb.DeviceArray.from_numpy(b))
    n0 = h.launch_count()
    sb.ldiv_(x, L, sb.DeviceArray.from_numpy(b))
    assert h.launch_count() - n0 == 1
    assert R.rel_err(x.numpy(), np.linalg.solve(A, b)) <= 1e-12


def test_time_dependent_matrix_refactorises(sb):
    rng = np.random.default_rng(4)
    n, k = 64, 2
    A0 = with_cond(rng, n, 1e8)

    def upd(A, u, p, t):
        A.t.mul_(0).add_(sb.DeviceArray.from_numpy(np.asfortranarray(A0 * (1.0 + t))).t)
    L = sb.factorize(sb.MatrixOperator(A0.copy(order="F"), update_func=upd))
    b = np.asfortranarray(rng.standard_normal((n, k)))
    sb.update_coefficients_(L, None, None, 2.0)
    x = sb.DeviceArray.full((n, k), float("nan"), np.float64)
    sb.ldiv_(x, L, sb.DeviceArray.from_numpy(b))
    assert backward_error(3.0 * A0, x.numpy(), b) <= 50 * EPS
