"""GPU parity tests for SURVEY.md §8(f) rank 4: BlockDiagonalOperator (src/block.jl:145-179) and WOperator
(src/woperator.jl:447-481) through the C ABI, against the oracle's sweep-by-sweep restatements and the dense
definitions the reference's tests use (test/block.jl, test/woperator.jl)."""
import numpy as np
import pytest
import scipy.linalg as sla

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


def rnd(rng, *shape, dtype=np.float64):
    return np.asfortranarray(rng.random(shape).astype(dtype))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("K", [1, 7, 64])
def test_block_diagonal_ragged(sb, dtype, K):
    rng = np.random.default_rng(21)
    d = (rng.random(4) + 0.5).astype(dtype)
    D = rnd(rng, 5, K, dtype=dtype)
    mats = [rnd(rng, 3, 5, dtype=dtype), np.diag(d), rnd(rng, 40, 24, dtype=dtype), np.eye(3, dtype=dtype),
            rnd(rng, 6, 2, dtype=dtype), np.zeros((2, 2), dtype)]
    ops = [sb.MatrixOperator(mats[0]), sb.DiagonalOperator(d), sb.MatrixOperator(np.asfortranarray(mats[2].T.copy()), trans=True),
           sb.IdentityOperator(3), sb.MatrixOperator(mats[4]), sb.NullOperator(2)]
    L = sb.BlockDiagonalOperator(*ops)
    dense = sla.block_diag(*[m.astype(np.float64) for m in mats])
    assert L.size() == dense.shape
    v = rnd(rng, dense.shape[1], K, dtype=dtype)
    w0 = rnd(rng, dense.shape[0], K, dtype=dtype)
    vd = sb.DeviceArray.from_numpy(v)
    L = sb.cache_operator(L, vd)
    w = sb.DeviceArray.full(w0.shape, float("nan"), dtype)
    sb.mul_(w, L, vd)
    want = R.block_diag_mul(np.empty_like(w0, dtype=np.float64), [m.astype(np.float64) for m in mats], v.astype(np.float64))
    assert R.rel_err(want, dense @ v) < 1e-12
    assert R.rel_err(w.numpy(), want) <= TOL[np.dtype(dtype)]
    w = sb.DeviceArray.from_numpy(w0)
    sb.mul_(w, L, vd, 0.7, 0.3)
    assert R.rel_err(w.numpy(), 0.7 * want + 0.3 * w0) <= TOL[np.dtype(dtype)]
    # batched-diagonal block needs the same K as v
    Lb = sb.BlockDiagonalOperator(sb.MatrixOperator(mats[0]), sb.DiagonalOperator(D))
    vb = rnd(rng, 10, K, dtype=dtype)
    wb = sb.DeviceArray.full((8, K), float("nan"), dtype)
    sb.mul_(wb, sb.cache_operator(Lb, sb.DeviceArray.from_numpy(vb)), sb.DeviceArray.from_numpy(vb))
    wantb = np.vstack([mats[0].astype(np.float64) @ vb[:5], D.astype(np.float64) * vb[5:]])
    assert R.rel_err(wb.numpy(), wantb) <= TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(8, 8, 5, 12), (128, 96, 6, 64), (33, 17, 3, 9), (256, 256, 16, 256)])
def test_block_diagonal_uniform_blocks_batched(sb, dtype, shape):
    """Equal dense blocks: one strided-batched GEMM (row slices of v / w as batch strides)."""
    m, n, nb, K = shape
    rng = np.random.default_rng(m + n + nb)
    mats = [rnd(rng, m, n, dtype=dtype) - dtype(0.5) for _ in range(nb)]
    L = sb.BlockDiagonalOperator(*mats)
    v = rnd(rng, n * nb, K, dtype=dtype)
    w0 = rnd(rng, m * nb, K, dtype=dtype)
    vd = sb.DeviceArray.from_numpy(v)
    L = sb.cache_operator(L, vd)
    h = sb.handle(0)
    n0 = h.launch_count()
    w = sb.DeviceArray.full(w0.shape, float("nan"), dtype)
    sb.mul_(w, L, vd)
    assert h.launch_count() - n0 == 1, "uniform blocks must run as one launch"
    want = np.vstack([mats[b].astype(np.float64) @ v[b * n:(b + 1) * n].astype(np.float64) for b in range(nb)])
    scale = np.linalg.norm(np.vstack([np.abs(mats[b].astype(np.float64)) @ np.abs(v[b * n:(b + 1) * n]) for b in range(nb)]))
    assert np.linalg.norm(w.numpy() - want) <= TOL[np.dtype(dtype)] * scale
    w = sb.DeviceArray.from_numpy(w0)
    sb.mul_(w, L, vd, 0.7, 0.3)
    assert np.linalg.norm(w.numpy() - (0.7 * want + 0.3 * w0)) <= TOL[np.dtype(dtype)] * (scale + np.linalg.norm(w0))
    # adjoint: blocks transposed
    La = sb.cache_operator(L.adjoint(), sb.DeviceArray.from_numpy(w0))
    u = sb.DeviceArray.full((n * nb, K), float("nan"), dtype)
    sb.mul_(u, La, sb.DeviceArray.from_numpy(w0))
    wantT = np.vstack([mats[b].astype(np.float64).T @ w0[b * m:(b + 1) * m].astype(np.float64) for b in range(nb)])
    assert R.rel_err(u.numpy(), wantT) <= 10 * TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,K", [(9, 1), (200, 1), (2048, 1), (300, 4), (512, 32)])
def test_woperator(sb, dtype, n, K):
    """`W = J - M/γ` (src/woperator.jl:447-481): dense and UniformScaling mass matrices, γ updated in place."""
    rng = np.random.default_rng(n + K)
    J, M = rnd(rng, n, n, dtype=dtype) - dtype(0.5), rnd(rng, n, n, dtype=dtype)
    gamma = 0.37
    shp = (n,) if K == 1 else (n, K)
    b = np.asfortranarray((rng.random(shp) - 0.5).astype(dtype))
    bd = sb.DeviceArray.from_numpy(b)
    f64 = lambda x: np.asarray(x, dtype=np.float64)
    tol = TOL[np.dtype(dtype)] * 4
    W = sb.cache_operator(sb.WOperator(M, gamma, J), bd)
    y = sb.DeviceArray.full(shp, float("nan"), dtype)
    sb.mul_(y, W, bd)
    want = R.woperator_mul(np.empty(shp), f64(M), gamma, f64(J), f64(b))
    scale = np.linalg.norm(np.abs(f64(J)) @ np.abs(f64(b)) + np.abs(f64(M)) @ np.abs(f64(b)) / gamma)
    assert np.linalg.norm(y.numpy() - want) <= tol * scale
    # new γ: nothing is re-planned, the scalar is read at call time (update_coefficients!(W, u, p, t; gamma))
    W.update_coefficients_(None, None, None, gamma=1.5)
    sb.mul_(y, W, bd)
    want = R.woperator_mul(np.empty(shp), f64(M), 1.5, f64(J), f64(b))
    assert np.linalg.norm(y.numpy() - want) <= tol * scale
    # UniformScaling mass matrix
    Wu = sb.cache_operator(sb.WOperator(2.0, gamma, J), bd)
    sb.mul_(y, Wu, bd)
    want = R.woperator_mul(np.empty(shp), 2.0, gamma, f64(J), f64(b))
    assert np.linalg.norm(y.numpy() - want) <= tol * scale


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_woperator_solve(sb, dtype):
    """`W \\ x` = `convert(AbstractMatrix, W) \\ x` (src/woperator.jl:431-445): the concrete form J - M/γ is assembled
    and inverted on the device once per (J, M, γ); a new γ re-forms it."""
    rng = np.random.default_rng(33)
    n, K = 96, 5
    J = np.asfortranarray((rng.random((n, n)) - 0.5).astype(dtype))
    M = np.asfortranarray((rng.random((n, n)) + n * np.eye(n)).astype(dtype))
    x = np.asfortranarray(rng.random((n, K)).astype(dtype))
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    for mass, Md in ((M, f64(M)), (2.0, 2.0 * np.eye(n))):
        W = sb.WOperator(mass, 0.37, J)
        assert sb.has_ldiv(W)
        for gamma in (0.37, 0.05):
            W.update_coefficients_(None, None, None, gamma=gamma)
            dense = f64(J) - Md / gamma
            want = np.linalg.solve(dense, f64(x))
            tol = TOL[np.dtype(dtype)] * max(1.0, np.linalg.cond(dense) / 100.0)
            y = sb.DeviceArray.full((n, K), float("nan"), dtype)
            sb.ldiv_(y, W, sb.DeviceArray.from_numpy(x))
            assert R.rel_err(y.numpy(), want) <= tol
            u = sb.DeviceArray.from_numpy(x)
            sb.ldiv_(W, u)
            assert R.rel_err(u.numpy(), want) <= tol
