"""GPU parity tests of the solve side: `factorize` + `ldiv!` through the C ABI against
the oracle's restatement of src/tensor.jl:612-673 / outer_div! and against the dense definition `AB \\ v`
(test/matrix.jl:646-671, :801-806).

Tolerance: a solve's forward error scales with the condition number, for LAPACK's getrs as much as for the
explicit inverse used here, so the bound is 1e-12 · max(1, κ/100) (Float64) and 1e-5 · max(1, κ/100) (Float32)
with κ computed from the dense factors; the test matrices (rand + n·I, like a mass/stiffness-type factor) have
κ < 10, where this is the plain north_star tolerance."""
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


def wellcond(rng, n, dtype=np.float64):
    return np.asfortranarray((rng.random((n, n)) + n * np.eye(n)).astype(dtype))


def tol(dtype, *mats):
    kappa = float(np.prod([np.linalg.cond(np.asarray(m, dtype=np.float64)) for m in mats]))
    return TOL[np.dtype(dtype)] * max(1.0, kappa / 100.0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 3, 17, 64, 130, 512])
def test_invert_matches_lapack(sb, dtype, n):
    rng = np.random.default_rng(n)
    A = np.asfortranarray((rng.random((n, n)) - 0.5).astype(dtype))      # general matrix: pivoting matters
    A[0, 0] = 0.0                                                        # ... and the first pivot must be exchanged
    if n == 1:
        A[0, 0] = 0.25
    got = sb.invert_matrix(A).numpy()
    want = np.linalg.inv(A.astype(np.float64))
    kappa = np.linalg.cond(A.astype(np.float64))
    eps = np.finfo(dtype).eps
    assert R.rel_err(got, want) <= 50 * kappa * eps, f"inverse error {R.rel_err(got, want):.3e}, cond {kappa:.2e}"
    resid = np.abs(A.astype(np.float64) @ got.astype(np.float64) - np.eye(n)).max()
    assert resid <= 50 * kappa * eps * max(1, n)


def test_invert_singular_raises(sb):
    A = np.asfortranarray(np.ones((8, 8)))
    with pytest.raises(sb.ScimlopError) as e:
        sb.invert_matrix(A)
    assert e.value.status == sb._lib.ERR_SINGULAR


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("vector_input", [False, True])
def test_tpo_ldiv_square_shapes(sb, dtype, vector_input):
    """test/matrix.jl:646-671 (K = 19) and :801-806 (vector input): 3-arg and 2-arg ldiv!, two and three factors."""
    rng = np.random.default_rng(11)
    n1, n2, n3 = 3, 5, 7
    A, B, Cm = wellcond(rng, n1, dtype), wellcond(rng, n2, dtype), wellcond(rng, n3, dtype)
    K = 1 if vector_input else 19
    shp = lambda n: (n,) if vector_input else (n, K)
    for mats in ([A, B], [A, B, Cm]):
        N = int(np.prod([m.shape[0] for m in mats]))
        v = np.asfortranarray(rng.random(shp(N)).astype(dtype))
        vd = sb.DeviceArray.from_numpy(v)
        L = sb.cache_operator(sb.factorize(sb.kron(*mats)), vd)
        assert sb.has_ldiv(L)
        w = sb.DeviceArray.full(shp(N), float("nan"), dtype)
        assert sb.ldiv_(w, L, vd) is w
        want = R.kron_solve([m.astype(np.float64) for m in mats], v.astype(np.float64))
        if len(mats) == 2:
            assert R.rel_err(R.tpo_ldiv(A.astype(np.float64), B.astype(np.float64), v.astype(np.float64)), want) < 1e-12
        assert R.rel_err(w.numpy(), want) <= tol(dtype, *mats)
        # mul! still works on the factorised operator and inverts ldiv!
        back = sb.DeviceArray.full(shp(N), float("nan"), dtype)
        sb.mul_(back, L, w)
        assert R.rel_err(back.numpy(), v) <= 10 * tol(dtype, *mats)
        # two-argument form
        u = sb.DeviceArray.from_numpy(v)
        assert sb.ldiv_(L, u) is u
        assert R.rel_err(u.numpy(), want) <= tol(dtype, *mats)


def test_tpo_ldiv_identity_scaled_diag_factors(sb):
    rng = np.random.default_rng(12)
    A, B = wellcond(rng, 6), wellcond(rng, 9)
    d = rng.random(9) + 0.5
    K = 7
    cases = [
        ([sb.IdentityOperator(6), sb.MatrixOperator(B)], [np.eye(6), B], 1.0),
        ([sb.MatrixOperator(A), sb.IdentityOperator(9)], [A, np.eye(9)], 1.0),
        ([2.0 * sb.MatrixOperator(A), sb.MatrixOperator(B)], [A, B], 2.0),
        ([sb.MatrixOperator(A), sb.DiagonalOperator(d)], [A, np.diag(d)], 1.0),
        ([sb.MatrixOperator(A).adjoint(), sb.MatrixOperator(B)], [A.T, B], 1.0),
    ]
    for ops, mats, lam in cases:
        v = np.asfortranarray(rng.random((54, K)))
        vd = sb.DeviceArray.from_numpy(v)
        L = sb.cache_operator(sb.factorize(sb.TensorProductOperator(*ops)), vd)
        w = sb.DeviceArray.full((54, K), float("nan"), np.float64)
        sb.ldiv_(w, L, vd)
        want = R.kron_solve(mats, v) / lam
        assert R.rel_err(w.numpy(), want) <= 1e-12
        u = sb.DeviceArray.from_numpy(v)          # in place, including the single-GEMM-stage cases
        sb.ldiv_(L, u)
        assert R.rel_err(u.numpy(), want) <= 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_leaf_and_tree_ldiv(sb, dtype):
    """Matrix (test/matrix.jl:84-90 style), Diagonal, BatchedDiagonal (src/batch.jl:181-200), Identity, Scaled
    (src/basic.jl:538-546), Composed (:1209-1256), InvertedOperator (:1420-1450)."""
    rng = np.random.default_rng(13)
    N, K = 33, 6
    A, B = wellcond(rng, N, dtype), wellcond(rng, N, dtype)
    d = (rng.random(N) + 0.5).astype(dtype)
    D = np.asfortranarray((rng.random((N, K)) + 0.5).astype(dtype))
    v = np.asfortranarray(rng.random((N, K)).astype(dtype))
    vd = sb.DeviceArray.from_numpy(v)
    f64 = lambda x: np.asarray(x, dtype=np.float64)

    def run(L, want, t):
        L = sb.cache_operator(L, vd)
        w = sb.DeviceArray.full((N, K), float("nan"), dtype)
        sb.ldiv_(w, L, vd)
        assert R.rel_err(w.numpy(), want) <= t
        u = sb.DeviceArray.from_numpy(v)
        sb.ldiv_(L, u)
        assert R.rel_err(u.numpy(), want) <= t
        return L

    t1 = tol(dtype, A)
    run(sb.factorize(sb.MatrixOperator(A)), np.linalg.solve(f64(A), f64(v)), t1)
    run(sb.DiagonalOperator(d), f64(v) / f64(d)[:, None], TOL[np.dtype(dtype)])
    run(sb.DiagonalOperator(D), f64(v) / f64(D), TOL[np.dtype(dtype)])
    run(sb.IdentityOperator(N), f64(v), 0.0)
    run(sb.factorize(3.0 * sb.MatrixOperator(A)), np.linalg.solve(f64(A), f64(v)) / 3.0, t1)
    comp = sb.factorize(sb.MatrixOperator(A) * sb.DiagonalOperator(d) * sb.MatrixOperator(B))
    want = np.linalg.solve(f64(B), np.linalg.solve(f64(A), f64(v)) / f64(d)[:, None])
    Lc = run(comp, want, tol(dtype, A, B))
    # mul! of the factorised composition is unchanged
    w = sb.DeviceArray.full((N, K), float("nan"), dtype)
    sb.mul_(w, Lc, vd)
    assert R.rel_err(w.numpy(), f64(A) @ (f64(d)[:, None] * (f64(B) @ f64(v)))) <= TOL[np.dtype(dtype)]
    # InvertedOperator: mul! is the solve, 3- and 5-arg
    Li = sb.cache_operator(sb.inv(sb.factorize(sb.MatrixOperator(A))), vd)
    w0 = np.asfortranarray(rng.random((N, K)).astype(dtype))
    w = sb.DeviceArray.from_numpy(w0)
    sb.mul_(w, Li, vd, 0.7, 0.3)
    assert R.rel_err(w.numpy(), 0.7 * np.linalg.solve(f64(A), f64(v)) + 0.3 * f64(w0)) <= t1
    sb.mul_(w, Li, vd)
    assert R.rel_err(w.numpy(), np.linalg.solve(f64(A), f64(v))) <= t1
    # an operator that was not factorised has no ldiv!
    with pytest.raises(TypeError):
        sb.ldiv_(w, sb.MatrixOperator(A), vd)
    assert not sb.has_ldiv(sb.MatrixOperator(A))


def test_factorised_operator_follows_update_coefficients(sb):
    """`update_coefficients!` on an InvertibleOperator updates L and F (src/matrix.jl:575-579)."""
    rng = np.random.default_rng(14)
    N, K = 24, 5
    A0 = wellcond(rng, N)
    v = np.asfortranarray(rng.random((N, K)))
    vd = sb.DeviceArray.from_numpy(v)

    def upd(Adev, u, p, t):      # A(t) = A0 * (1 + t): rewritten in place on the device
        Adev.t.copy_(sb.DeviceArray.from_numpy(np.asfortranarray(A0 * (1.0 + t))).t)

    L = sb.cache_operator(sb.factorize(sb.MatrixOperator(A0.copy(order="F"), update_func=upd)), vd)
    w = sb.DeviceArray.full((N, K), float("nan"), np.float64)
    sb.update_coefficients_(L, None, None, 1.0)
    sb.ldiv_(w, L, vd)
    assert R.rel_err(w.numpy(), np.linalg.solve(2.0 * A0, v)) <= 1e-12


def test_config2_size_ldiv_round_trip(sb):
    """Headline shape: 512x512 (x) 512x512 Float64, 64 columns: ldiv!(mul!(v)) == v, and sampled columns against
    the LU-based oracle."""
    import torch
    rng = np.random.default_rng(15)
    A, B = wellcond(rng, 512), wellcond(rng, 512)
    k = 64
    g = torch.Generator(device="cuda").manual_seed(7)
    V = sb.DeviceArray(torch.rand(512 * 512 * k, dtype=torch.float64, device="cuda", generator=g), (512 * 512, k))
    L = sb.cache_operator(sb.factorize(sb.kron(A, B)), V)
    W = sb.DeviceArray.empty((512 * 512, k), np.float64)
    X = sb.DeviceArray.empty((512 * 512, k), np.float64)
    sb.mul_(W, L, V)
    sb.ldiv_(X, L, W)
    rel = float(((X.t - V.t).norm() / V.t.norm()).item())
    assert rel <= 1e-12, rel
    cols = [0, 31, 63]
    Wh = np.asfortranarray(W.numpy()[:, cols])
    want = R.kron_solve([A, B], Wh)
    assert R.rel_err(X.numpy()[:, cols], want) <= 1e-12
