"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI, against the CPU
oracle (oracle/) on identical seeded inputs.

Mirrors the reference's own tests: every check is "lazy operator ≈ dense
materialisation × v" (test/matrix.jl:625-644, test/basic.jl:365-384), plus the exact known answers of
test/precompile_workload.jl.  Tolerances are the north_star's: rel. error <= 1e-12 (Float64), <= 1e-5
(Float32), norm-wise metric of `isapprox` (oracle.ref_numpy.rel_err)."""
import ctypes as C

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
    sb.handle(0)  # fails loudly if the extension is missing or the device is not sm_100
    return sb


def rnd(rng, *shape, dtype=np.float64):
    return np.asfortranarray(rng.random(shape).astype(dtype))


def dev(sb, a):
    return sb.DeviceArray.from_numpy(a)


def check(got, want, dtype, what=""):
    err = R.rel_err(got, want)
    assert np.isfinite(np.asarray(got, dtype=np.float64)).all(), f"{what}: non-finite output"
    assert err <= TOL[np.dtype(dtype)], f"{what}: rel err {err:.3e} > {TOL[np.dtype(dtype)]:.0e}"


# ---- GEMM engine (K1-K9): both the TMA/DMMA path and the shape-agnostic path --------------------------
GEMM_CASES = [
    # M, N, K, batch, transA, transB
    (128, 128, 16, 1, 0, 0), (128, 128, 64, 1, 0, 1), (128, 128, 64, 1, 1, 0), (128, 128, 64, 1, 1, 1),
    (256, 384, 128, 1, 0, 0), (100, 200, 100, 3, 0, 1), (130, 70, 36, 2, 1, 0), (64, 64, 512, 4, 0, 0),
    (512, 512, 512, 2, 0, 1), (18, 22, 6, 5, 1, 1), (300, 40, 20, 1, 0, 0),
    (3, 5, 7, 1, 0, 0), (7, 11, 13, 4, 0, 1), (13, 17, 19, 2, 1, 0), (33, 1, 9, 1, 0, 0), (1, 1, 1, 1, 1, 1),
    (65, 67, 1, 3, 0, 0),
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", GEMM_CASES)
def test_gemm_strided_batched(sb, case, dtype):
    M, N, K, batch, ta, tb = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    A = rnd(rng, *((K, M) if ta else (M, K)), batch, dtype=dtype)
    B = rnd(rng, *((N, K) if tb else (K, N)), batch, dtype=dtype)
    C0 = rnd(rng, M, N, batch, dtype=dtype)
    alpha, beta = 0.7, 0.3
    h = sb.handle(0)
    import torch
    for (a, b) in ((alpha, beta), (alpha, 0.0), (None, None)):
        dA, dB = dev(sb, A), dev(sb, B)
        dC = dev(sb, C0 if (b not in (None, 0.0)) else np.full_like(C0, np.nan))
        ct = C.c_double if dtype == np.float64 else C.c_float
        pa = C.byref(ct(a)) if a is not None else None
        pb = C.byref(ct(b)) if b is not None else None
        rc = h.lib.scimlop_gemm_strided_batched(
            h.ptr, 0 if dtype == np.float64 else 1, 0, ta, tb, M, N, K, pa, C.c_void_p(dA.ptr), A.shape[0],
            A.shape[0] * A.shape[1], C.c_void_p(dB.ptr), B.shape[0], B.shape[0] * B.shape[1], pb,
            C.c_void_p(dC.ptr), M, M * N, batch, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        h.check(rc, "gemm")
        got = dC.numpy()
        want = np.empty((M, N, batch))
        for i in range(batch):
            opA = A[:, :, i].T if ta else A[:, :, i]
            opB = B[:, :, i].T if tb else B[:, :, i]
            prod = opA.astype(np.float64) @ opB.astype(np.float64)
            want[:, :, i] = (1.0 if a is None else a) * prod + (0.0 if not b else b * C0[:, :, i])
        check(got, want, dtype, f"gemm {case} alpha={a} beta={b}")


TF32_CASES = [
    # M, N, K, batch, transA, transB  -- shapes that are eligible for the tcgen05 (3xTF32) path
    (128, 4096, 128, 1, 0, 0),    # (a) factor = A, K-major X, MN-major factor     [innermost Kronecker mode]
    (128, 4096, 128, 1, 1, 0),    # (a) transposed factor -> K-major factor
    (100, 1000, 72, 1, 0, 0),     # (a) ragged: m, n, rows not multiples of the tile
    (48, 300, 12, 1, 1, 0),       # (a) tiny factor, K tail
    (128, 128, 128, 64, 0, 1),    # (b) factor = B shared, MN-major X              [outer Kronecker modes]
    (16384, 128, 128, 2, 0, 1),   # (b) many row tiles per batch
    (300, 40, 20, 3, 0, 0),       # (b) K-major factor, ragged rows
    (1000, 100, 72, 5, 0, 1),     # (b) ragged everything
]


@pytest.mark.parametrize("case", TF32_CASES)
@pytest.mark.parametrize("precision,tol", [(0, 1e-5), (1, 1e-5), (3, 1e-5), (2, 3e-3)])
def test_gemm_f32_tensor_core_paths(sb, case, precision, tol):
    """Float32: DEFAULT / TF32X3 run the tcgen05 3xTF32 kernel and must hold 1e-5; FP32_EXACT forces FFMA;
    single-pass TF32 is opt-in with its own (documented) tolerance."""
    import torch
    M, N, K, batch, ta, tb = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    A = rnd(rng, *((K, M) if ta else (M, K)), batch, dtype=np.float32)
    shared_b = batch > 1          # path (b) needs the factor shared across the batch
    B = rnd(rng, *((N, K) if tb else (K, N)), 1 if shared_b else batch, dtype=np.float32)
    C0 = rnd(rng, M, N, batch, dtype=np.float32)
    h = sb.handle(0)
    for (a, b) in ((0.7, 0.3), (None, None)):
        dA, dB = dev(sb, A), dev(sb, B)
        dC = dev(sb, C0 if b else np.full_like(C0, np.nan))
        pa = C.byref(C.c_float(a)) if a is not None else None
        pb = C.byref(C.c_float(b)) if b is not None else None
        n0 = h.launch_count()
        rc = h.lib.scimlop_gemm_strided_batched(
            h.ptr, 1, precision, ta, tb, M, N, K, pa, C.c_void_p(dA.ptr), A.shape[0], A.shape[0] * A.shape[1],
            C.c_void_p(dB.ptr), B.shape[0], 0 if shared_b else B.shape[0] * B.shape[1], pb, C.c_void_p(dC.ptr), M, M * N,
            batch, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        h.check(rc, "gemm f32")
        assert h.launch_count() == n0 + 1
        want = np.empty((M, N, batch))
        for i in range(batch):
            opA = (A[:, :, i].T if ta else A[:, :, i]).astype(np.float64)
            Bi = B[:, :, 0 if shared_b else i]
            opB = (Bi.T if tb else Bi).astype(np.float64)
            want[:, :, i] = (1.0 if a is None else a) * (opA @ opB) + (b * C0[:, :, i] if b else 0.0)
        err = R.rel_err(dC.numpy(), want)
        assert np.isfinite(dC.numpy()).all()
        assert err <= tol, f"{case} precision={precision} alpha={a}: rel err {err:.3e}"


def test_gemm_f32_tensor_core_large_repeated(sb):
    """Many tiles per CTA, launched repeatedly: the mbarrier rings of the tcgen05 kernel must stay in phase (a
    ring depth that is not a multiple of the splitter-group count faulted here).  At this size the check is a
    size-independent property: the tensor-core path and the exact-FP32 FFMA path (itself oracle-checked above)
    agree to 1e-5, and a sampled block agrees with the Float64 oracle product."""
    import torch
    h = sb.handle(0)
    M, N, K = 128, 524288, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.rand(M * K, device="cuda", generator=g)
    B = torch.rand(K * N, device="cuda", generator=g)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    outs = {}
    for prec, reps in ((3, 1), (0, 4)):
        for rep in range(reps):
            Cc = torch.full((M * N,), float("nan"), device="cuda")
            rc = h.lib.scimlop_gemm_strided_batched(h.ptr, 1, prec, 0, 0, M, N, K, None, C.c_void_p(A.data_ptr()), M, 0,
                                                    C.c_void_p(B.data_ptr()), K, 0, None, C.c_void_p(Cc.data_ptr()), M, 0, 1, st)
            h.check(rc, "gemm f32 large")
            torch.cuda.synchronize()
            outs[(prec, rep)] = Cc
    ref = outs[(3, 0)]
    for rep in range(4):
        err = float((outs[(0, rep)] - ref).norm() / ref.norm())
        assert err <= 1e-5, f"rep {rep}: tensor-core vs FFMA rel err {err:.3e}"
    Ah = A.cpu().numpy().reshape((M, K), order="F").astype(np.float64)
    Bh = B[: K * 64].cpu().numpy().reshape((K, 64), order="F").astype(np.float64)
    check(outs[(0, 3)][: M * 64].cpu().numpy().reshape((M, 64), order="F"), Ah @ Bh, np.float32, "sampled block vs oracle")


@pytest.mark.parametrize("profile", ["small", "large"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gemm_dispatch_fuzz(sb, dtype, profile):
    """Seeded fuzz over the GEMM dispatcher: random extents (including 1 and non-multiples of every tile),
    padded leading dimensions (even/odd), offset base pointers, shared or batched operands, all transposes,
    alpha/beta on or off -- whatever kernel the dispatcher picks must agree with the Float64 product.
    "large" draws the shapes of the few-column (skinny), split-K, and streamed tcgen05 paths and their boundaries."""
    import torch
    h = sb.handle(0)
    rng = np.random.default_rng((1234 if dtype == np.float64 else 4321) + (7 if profile == "large" else 0))
    es = np.dtype(dtype).itemsize
    ct = C.c_double if dtype == np.float64 else C.c_float
    dt = 0 if dtype == np.float64 else 1
    sizes = [1, 2, 5, 16, 31, 32, 33, 64, 100, 128, 129, 200, 256, 300]
    for trial in range(40 if profile == "small" else 30):
        M, N, K = (int(rng.choice(sizes)) for _ in range(3))
        if trial % 5 == 0:
            N = int(rng.choice([1024, 2048, 4100]))      # long streamed side: tensor-core eligible shapes
        batch = int(rng.choice([1, 1, 2, 5]))
        if profile == "large":
            M = int(rng.choice([16, 64, 128, 130, 512, 700, 1024, 2048]))
            K = int(rng.choice([32, 63, 64, 65, 128, 129, 512, 1000, 2048]))
            N = int(rng.choice([1, 2, 3, 8, 9, 15, 16, 17, 48, 49, 64, 65, 130, 512]))
            batch = int(rng.choice([1, 1, 1, 3]))
        ta, tb = int(rng.integers(2)), int(rng.integers(2))
        ar, ac = (K, M) if ta else (M, K)
        br, bc = (N, K) if tb else (K, N)
        lda, ldb, ldc = ar + int(rng.choice([0, 0, 1, 2, 4])), br + int(rng.choice([0, 0, 1, 2, 4])), M + int(rng.choice([0, 0, 1, 4]))
        offa, offb = int(rng.choice([0, 0, 1, 2, 4])), int(rng.choice([0, 0, 1, 2, 4]))
        shared_b = bool(rng.integers(2)) and batch > 1
        sA, sB, sC = lda * ac, (0 if shared_b else ldb * bc), ldc * N
        Abuf = rnd(rng, offa + sA * batch, dtype=dtype)
        Bbuf = rnd(rng, offb + max(sB, ldb * bc) * (1 if shared_b else batch), dtype=dtype)
        C0 = rnd(rng, sC * batch, dtype=dtype)
        use_ab = bool(rng.integers(2))
        a, b = (float(rng.random()) + 0.5, float(rng.choice([0.0, 0.3]))) if use_ab else (None, None)
        dA, dB = dev(sb, Abuf), dev(sb, Bbuf)
        dC = dev(sb, C0 if b else np.where(np.arange(C0.size) % 3 == 0, np.nan, C0).astype(dtype))
        rc = h.lib.scimlop_gemm_strided_batched(
            h.ptr, dt, 0, ta, tb, M, N, K, C.byref(ct(a)) if use_ab else None, C.c_void_p(dA.ptr + es * offa), lda, sA,
            C.c_void_p(dB.ptr + es * offb), ldb, sB, C.byref(ct(b)) if use_ab else None, C.c_void_p(dC.ptr), ldc, sC, batch,
            C.c_void_p(torch.cuda.current_stream().cuda_stream))
        h.check(rc, "gemm fuzz")
        got = dC.numpy()
        for i in range(batch):
            Am = Abuf[offa + i * sA: offa + (i + 1) * sA].reshape((lda, ac), order="F")[:ar].astype(np.float64)
            bo = offb + (0 if shared_b else i * sB)
            Bm = Bbuf[bo: bo + ldb * bc].reshape((ldb, bc), order="F")[:br].astype(np.float64)
            Cm0 = C0[i * sC:(i + 1) * sC].reshape((ldc, N), order="F")[:M].astype(np.float64)
            want = (1.0 if a is None else a) * ((Am.T if ta else Am) @ (Bm.T if tb else Bm)) + (b * Cm0 if b else 0.0)
            gi = got[i * sC:(i + 1) * sC].reshape((ldc, N), order="F")
            check(gi[:M], want, dtype, f"trial {trial}: M{M} N{N} K{K} b{batch} t{ta}{tb} ld{lda},{ldb},{ldc} off{offa},{offb} ab={a},{b}")
            if ldc > M and b:   # padding rows of C are never touched
                assert (gi[M:] == C0[i * sC:(i + 1) * sC].reshape((ldc, N), order="F")[M:]).all()


def test_gemm_shared_operand_and_odd_ld(sb):
    """stride 0 = shared operand; odd leading dimensions / unaligned bases take the shape-agnostic kernel."""
    rng = np.random.default_rng(5)
    import torch
    h = sb.handle(0)
    M, N, K, batch = 96, 80, 48, 3
    for lda_pad, off in ((0, 0), (1, 0), (0, 1), (3, 1)):
        Abuf = rnd(rng, (M + lda_pad) * K * batch + 4)
        Bbuf = rnd(rng, K * N + 4)
        Cbuf = np.zeros(M * N * batch)
        dA, dB, dC = dev(sb, Abuf), dev(sb, Bbuf), dev(sb, Cbuf)
        lda = M + lda_pad
        rc = h.lib.scimlop_gemm_strided_batched(
            h.ptr, 0, 0, 0, 0, M, N, K, None, C.c_void_p(dA.ptr + 8 * off), lda, lda * K, C.c_void_p(dB.ptr + 8 * off),
            K, 0, None, C.c_void_p(dC.ptr), M, M * N, batch, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        h.check(rc, "gemm")
        got = dC.numpy().reshape((M, N, batch), order="F")
        Bm = Bbuf[off:off + K * N].reshape((K, N), order="F")
        for i in range(batch):
            Am = Abuf[off + i * lda * K: off + (i + 1) * lda * K].reshape((lda, K), order="F")[:M]
            check(got[:, :, i], Am @ Bm, np.float64, f"lda_pad={lda_pad} off={off} batch {i}")


# ---- exact known answers: test/precompile_workload.jl:10-37 ---------------------------------------------
def test_known_answers_exact(sb):
    v = dev(sb, np.array([1.0, 2.0]))
    A = sb.MatrixOperator(np.array([[2.0, 1.0], [1.0, 3.0]]))
    D = sb.DiagonalOperator(np.array([2.0, 3.0]))
    assert (A * v).numpy().tolist() == [4.0, 7.0]                                   # :10-14
    assert ((A + D) * v).numpy().tolist() == [6.0, 13.0]                            # :22
    AD = sb.cache_operator(A * D, v)
    w = sb.DeviceArray.zeros(2, np.float64)
    assert sb.mul_(w, AD, v).numpy().tolist() == [10.0, 20.0]                       # :23, :33-37
    assert ((2.0 * A) * v).numpy().tolist() == [8.0, 14.0]                          # :25
    v4 = dev(sb, np.array([1.0, 2.0, 3.0, 4.0]))
    T = sb.cache_operator(sb.kron(sb.IdentityOperator(2), D), v4)
    w4 = sb.DeviceArray.zeros(4, np.float64)
    assert sb.mul_(w4, T, v4).numpy().tolist() == [2.0, 6.0, 6.0, 12.0]             # :30-31


# ---- TPO shape matrix: test/matrix.jl:510-517, K=19; 3-arg :625-631, 5-arg :637-644, vector :793-799 -----
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("square", [False, True])
@pytest.mark.parametrize("vector_input", [False, True])
def test_tpo_matches_dense_kron(sb, dtype, square, vector_input):
    rng = np.random.default_rng(0)
    m1, n1, m2, n2, m3, n3 = 3, 5, 7, 11, 13, 17
    if square:
        n1, n2, n3 = m1, m2, m3
    K = 19
    A, B, Cm = rnd(rng, m1, n1, dtype=dtype), rnd(rng, m2, n2, dtype=dtype), rnd(rng, m3, n3, dtype=dtype)
    alpha, beta = float(rng.random()), float(rng.random())
    for mats in ((A, B), (A, B, Cm)):
        dense = mats[0].astype(np.float64)
        for M in mats[1:]:
            dense = np.kron(dense, M.astype(np.float64))
        m, n = dense.shape
        v = rnd(rng, n, dtype=dtype) if vector_input else rnd(rng, n, K, dtype=dtype)
        w0 = rnd(rng, *((m,) + v.shape[1:]), dtype=dtype)
        dv = dev(sb, v)
        L = sb.TensorProductOperator(*mats)
        assert L.size() == (m, n)
        with pytest.raises(AssertionError):          # @assert iscached(L), src/tensor.jl:548
            sb.mul_(dev(sb, w0), L, dv)
        L = sb.cache_operator(L, dv)
        w = dev(sb, np.full_like(w0, np.nan))        # 3-arg must not read w
        check(sb.mul_(w, L, dv).numpy(), dense @ v, dtype, "3-arg")
        w = dev(sb, w0)
        check(sb.mul_(w, L, dv, alpha, beta).numpy(), alpha * (dense @ v) + beta * w0, dtype, "5-arg")
        w = dev(sb, w0)                              # call form L(w, v, u, p, t, α, β), src/tensor.jl:951-957
        check(L(w, dv, None, None, 0.0, alpha, beta).numpy(), alpha * (dense @ v) + beta * w0, dtype, "call form")
        check((L * dv).numpy(), dense @ v, dtype, "out of place")
        # the restated reference algorithms agree too (stdlib permute path and LV path)
        for lv in (False, True):
            ref = R.TensorProductOperator(*[np.asarray(x, dtype=np.float64) for x in mats], lv=lv)
            v64 = np.asfortranarray(v.astype(np.float64))
            ref = ref.cache_operator(v64)
            wr = np.array(w0, dtype=np.float64, order="F", copy=True)
            ref.mul5(wr, v64, alpha, beta)
            w = dev(sb, w0)
            check(sb.mul_(w, L, dv, alpha, beta).numpy(), wr, dtype, f"vs restated reference lv={lv}")


# ---- identity placements: test/matrix.jl:696-703, :755-758 ------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_tpo_identity_and_diagonal_factors(sb, dtype):
    rng = np.random.default_rng(1)
    A, B = rnd(rng, 6, 6, dtype=dtype), rnd(rng, 10, 10, dtype=dtype)
    d = rnd(rng, 4, dtype=dtype)
    K = 12
    I4, I3 = np.eye(4), np.eye(3)
    cases = [
        ([("I", 4), ("M", A)], [I4, A]), ([("M", A), ("I", 4)], [A, I4]),
        ([("I", 3), ("M", A), ("I", 4)], [I3, A, I4]), ([("M", A), ("I", 4), ("M", B)], [A, I4, B]),
        ([("I", 3), ("I", 4)], [I3, I4]), ([("D", d), ("M", A)], [np.diag(d), A]),
        ([("M", A), ("D", d), ("I", 3)], [A, np.diag(d), I3]), ([("I", 3), ("D", d)], [I3, np.diag(d)]),
        ([("M", A), ("M", B), ("D", d)], [A, B, np.diag(d)]),
    ]
    for spec, mats in cases:
        ops = []
        for kind, x in spec:
            ops.append(sb.IdentityOperator(x) if kind == "I" else sb.DiagonalOperator(x) if kind == "D"
                       else sb.MatrixOperator(x))
        dense = np.asarray(mats[0], dtype=np.float64)
        for M in mats[1:]:
            dense = np.kron(dense, np.asarray(M, dtype=np.float64))
        n = dense.shape[1]
        for v in (rnd(rng, n, K, dtype=dtype), rnd(rng, n, dtype=dtype)):
            w0 = rnd(rng, *v.shape, dtype=dtype)
            dv = dev(sb, v)
            L = sb.cache_operator(sb.TensorProductOperator(*ops), dv)
            check(sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv).numpy(), dense @ v, dtype, f"{[s[0] for s in spec]} 3-arg")
            check(sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy(), 0.7 * (dense @ v) + 0.3 * w0, dtype,
                  f"{[s[0] for s in spec]} 5-arg")


def test_tpo_scaled_and_adjoint_factors(sb):
    """ScaledOperator outer (the reference's broken dispatch, src/tensor.jl:800-804: intended = scale by λ)
    and adjoint factors (TT' of test/total.jl:117-130)."""
    rng = np.random.default_rng(2)
    A, B = rnd(rng, 5, 8), rnd(rng, 9, 4)
    K = 6
    v = rnd(rng, 8 * 4, K)
    dv = dev(sb, v)
    L = sb.cache_operator(sb.kron(2.5 * sb.MatrixOperator(A), sb.MatrixOperator(B)), dv)
    w0 = rnd(rng, 45, K)
    check(sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy(), 0.7 * 2.5 * (np.kron(A, B) @ v) + 0.3 * w0, np.float64, "scaled outer")
    Lt = sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)).adjoint()
    vt = rnd(rng, 45, K)
    dvt = dev(sb, vt)
    Lt = sb.cache_operator(Lt, dvt)
    assert Lt.size() == (32, 45)
    check((Lt * dvt).numpy(), np.kron(A, B).T @ vt, np.float64, "adjoint TPO")
    w0 = rnd(rng, 32, K)
    check(sb.mul_(dev(sb, w0), Lt, dvt, 0.7, 0.3).numpy(), 0.7 * (np.kron(A, B).T @ vt) + 0.3 * w0, np.float64, "adjoint TPO 5-arg")


def test_tpo_nonnative_inner_uses_outer_hook(sb, oracle_c):
    """inner factor is an AddedOperator -> two-stage reference algorithm with the native
    `tensor_outer_mul_fast!` replacement (src/tensor.jl:773-776, :816-819)."""
    rng = np.random.default_rng(3)
    A, B1, B2 = rnd(rng, 6, 7), rnd(rng, 8, 8), rnd(rng, 8, 8)
    K = 5
    v = rnd(rng, 7 * 8, K)
    dv = dev(sb, v)
    L = sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B1) + sb.MatrixOperator(B2))
    L = sb.cache_operator(L, dv)
    assert L._native is None
    dense = np.kron(A, B1 + B2)
    w0 = rnd(rng, 48, K)
    check(sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv).numpy(), dense @ v, np.float64, "3-arg")
    check(sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy(), 0.7 * (dense @ v) + 0.3 * w0, np.float64, "5-arg")
    # the hook itself against the C restatement of the LoopVectorization nest
    import torch
    h = sb.handle(0)
    mi, mo, no = 8, 6, 7
    Cm = rnd(rng, mi * no, K)
    W0 = rnd(rng, mi * mo, K)
    want = oracle_c.lv_outer_mul(W0.copy(order="F"), A, Cm, mi, mo, no, K, 0.7, 0.3)
    dW, dA, dC = dev(sb, W0), dev(sb, A), dev(sb, Cm)
    rc = h.lib.scimlop_outer_mul_fast(h.ptr, 0, 0, C.c_void_p(dW.ptr), C.c_void_p(dA.ptr), mo, C.c_void_p(dC.ptr), mi, mo,
                                      no, K, C.byref(C.c_double(0.7)), C.byref(C.c_double(0.3)),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
    h.check(rc, "outer_mul_fast")
    check(dW.numpy(), want, np.float64, "outer_mul_fast vs LV restatement")


# ---- TensorSum: test/matrix.jl:809-867 -------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_tensor_sum(sb, dtype):
    rng = np.random.default_rng(4)
    for (mo, mi, K) in ((5, 7, 9), (32, 48, 3), (128, 128, 2)):
        A, B = rnd(rng, mo, mo, dtype=dtype), rnd(rng, mi, mi, dtype=dtype)
        dense = np.kron(A.astype(np.float64), np.eye(mi)) + np.kron(np.eye(mo), B.astype(np.float64))
        v = rnd(rng, mo * mi, K, dtype=dtype)
        w0 = rnd(rng, mo * mi, K, dtype=dtype)
        dv = dev(sb, v)
        L = sb.cache_operator(sb.kronsum(A, B), dv)
        check(sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv).numpy(), dense @ v, dtype, "kronsum 3-arg")
        check(sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy(), 0.7 * (dense @ v) + 0.3 * w0, dtype, "kronsum 5-arg")
        # the same thing spelled as an AddedOperator of two TPOs (config 5)
        L2 = sb.kron(sb.MatrixOperator(A), sb.IdentityOperator(mi)) + sb.kron(sb.IdentityOperator(mo), sb.MatrixOperator(B))
        L2 = sb.cache_operator(L2, dv)
        check(sb.mul_(dev(sb, w0), L2, dv, 0.7, 0.3).numpy(), 0.7 * (dense @ v) + 0.3 * w0, dtype, "Added of TPOs")


# ---- leaves: test/matrix.jl:72-80 (Matrix), :250-259 (Diagonal), :289-298 (BatchedDiagonal), basic.jl ----
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("N,K", [(8, 19), (8, 1), (64, 12), (1001, 7), (4096, 64)])
def test_leaves(sb, dtype, N, K):
    rng = np.random.default_rng(N * 31 + K)
    A, d, D = rnd(rng, N, N, dtype=dtype), rnd(rng, N, dtype=dtype), rnd(rng, N, K, dtype=dtype)
    v, w0 = rnd(rng, N, K, dtype=dtype), rnd(rng, N, K, dtype=dtype)
    dv = dev(sb, v)
    a, b = 0.7, 0.3
    A64, v64 = A.astype(np.float64), v.astype(np.float64)
    ops = [
        (sb.MatrixOperator(A), A64 @ v64),
        (sb.MatrixOperator(A).adjoint(), A64.T @ v64),
        (sb.DiagonalOperator(d), d[:, None].astype(np.float64) * v64),
        (sb.DiagonalOperator(D), D.astype(np.float64) * v64),
        (sb.IdentityOperator(N), v64),
        (sb.NullOperator(N), np.zeros_like(v64)),
    ]
    for L, want in ops:
        name = type(L).__name__
        check(sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv).numpy(), want, dtype, f"{name} 3-arg")
        check(sb.mul_(dev(sb, w0), L, dv, a, b).numpy(), a * want + b * w0, dtype, f"{name} 5-arg")
        # β == 0 must not read w: NaN may not leak (src/batch.jl:222-223, src/basic.jl:251,521)
        check(sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv, a, 0.0).numpy(), a * want, dtype, f"{name} beta=0")
        check(L(dev(sb, w0), dv, None, None, 0.0, a, b).numpy(), a * want + b * w0, dtype, f"{name} call form")


# ---- trees: test/basic.jl:258-318 (Scaled), :320-427 (Added), :516-608 (Composed), README.md:69-73 --------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("N,K", [(8, 12), (96, 5), (257, 3), (300, 40), (1000, 24)])   # the last two: split-K DMMA tiles with diagonal epilogues
def test_operator_trees(sb, dtype, N, K):
    rng = np.random.default_rng(N + K)
    A, B, Cm = (rnd(rng, N, N, dtype=dtype) for _ in range(3))
    d1, d2 = rnd(rng, N, dtype=dtype), rnd(rng, N, dtype=dtype)
    Dk = rnd(rng, N, K, dtype=dtype)
    v, w0 = rnd(rng, N, K, dtype=dtype), rnd(rng, N, K, dtype=dtype)
    f = lambda x: np.asarray(x, dtype=np.float64)
    I = np.eye(N)
    M, Mo = sb.MatrixOperator, lambda x: R.MatrixOperator(f(x))
    cases = {
        "A+B": (M(A) + M(B), f(A) + f(B)),
        "A-B": (M(A) - M(B), f(A) - f(B)),
        "2A": (2.0 * M(A), 2 * f(A)),
        "A*3": (M(A) * 3.0, 3 * f(A)),
        "A/4": (M(A) / 4.0, f(A) / 4),
        "-A": (-M(A), -f(A)),
        "A*B": (M(A) * M(B), f(A) @ f(B)),
        "A*B*C": (M(A) * M(B) * M(Cm), f(A) @ f(B) @ f(Cm)),
        "A+I": (M(A) + sb.IdentityOperator(N), f(A) + I),
        "A+2": (M(A) + 2.0, f(A) + 2 * I),
        "D*A'": (sb.DiagonalOperator(d1) * M(A).adjoint(), np.diag(f(d1)) @ f(A).T),
        "A*D": (M(A) * sb.DiagonalOperator(d1), f(A) @ np.diag(f(d1))),
        "D1*A*D2*B": (sb.DiagonalOperator(d1) * M(A) * sb.DiagonalOperator(d2) * M(B),
                      np.diag(f(d1)) @ f(A) @ np.diag(f(d2)) @ f(B)),
        "README D*M'+2M+I (config 4)": (sb.DiagonalOperator(d1) * M(A).adjoint() + 2.0 * M(A) + sb.IdentityOperator(N),
                                        np.diag(f(d1)) @ f(A).T + 2 * f(A) + I),
        "D1*D2 + 3I - D1": (sb.DiagonalOperator(d1) * sb.DiagonalOperator(d2) + 3.0 * sb.IdentityOperator(N)
                            - sb.DiagonalOperator(d1), np.diag(f(d1) * f(d2)) + 3 * I - np.diag(f(d1))),
        "six diagonal terms (two elementwise passes)": (
            sb.DiagonalOperator(d1) + sb.DiagonalOperator(d2) + 2.0 * sb.DiagonalOperator(d1) * sb.DiagonalOperator(d2)
            + sb.IdentityOperator(N) - 0.5 * sb.DiagonalOperator(d2) + 3.0 * sb.IdentityOperator(N),
            np.diag(f(d1) + f(d2) + 2 * f(d1) * f(d2) - 0.5 * f(d2)) + 4 * I),
        "0*A + B": (0.0 * M(A) + M(B), f(B)),
        "A + Null": (M(A) + sb.NullOperator(N), f(A)),
        "Null*A": (sb.NullOperator(N) * M(A), np.zeros((N, N))),
        "(A+B)*C (sum inside product)": ((M(A) + M(B)) * M(Cm), (f(A) + f(B)) @ f(Cm)),
    }
    dv = dev(sb, v)
    for name, (L, dense) in cases.items():
        L = sb.cache_operator(L, dv)
        want = dense @ f(v)
        check(sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv).numpy(), want, dtype, f"{name} 3-arg")
        check(sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy(), 0.7 * want + 0.3 * f(w0), dtype, f"{name} 5-arg")
        check(sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv, 0.7, 0.0).numpy(), 0.7 * want, dtype, f"{name} beta=0")
    # batched-diagonal trees: one dense matrix per column (oracle.dense_apply)
    bd_cases = {
        "Dk": (sb.DiagonalOperator(Dk), R.BatchedDiagonalOperator(f(Dk))),
        "Dk*A": (sb.DiagonalOperator(Dk) * M(A), R.ComposedOperator(R.BatchedDiagonalOperator(f(Dk)), Mo(A))),
        "A*Dk": (M(A) * sb.DiagonalOperator(Dk), R.ComposedOperator(Mo(A), R.BatchedDiagonalOperator(f(Dk)))),
        "config-4-elt: D1*Dk + 2*D2 + I": (
            sb.DiagonalOperator(d1) * sb.DiagonalOperator(Dk) + 2.0 * sb.DiagonalOperator(d2) + sb.IdentityOperator(N),
            R.AddedOperator(R.ComposedOperator(R.DiagonalOperator(f(d1)), R.BatchedDiagonalOperator(f(Dk))),
                            R.ScaledOperator(2.0, R.DiagonalOperator(f(d2))), R.IdentityOperator(N))),
        "Dk*A' + 2A + I": (
            sb.DiagonalOperator(Dk) * M(A).adjoint() + 2.0 * M(A) + sb.IdentityOperator(N),
            R.AddedOperator(R.ComposedOperator(R.BatchedDiagonalOperator(f(Dk)), Mo(A).adjoint()),
                            R.ScaledOperator(2.0, Mo(A)), R.IdentityOperator(N))),
    }
    for name, (L, Lref) in bd_cases.items():
        L = sb.cache_operator(L, dv)
        want = R.dense_apply(Lref, f(v))
        check(sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv).numpy(), want, dtype, f"{name} 3-arg")
        check(sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy(), 0.7 * want + 0.3 * f(w0), dtype, f"{name} 5-arg")
        # and the literal sequential evaluation of the reference (sweep by sweep) agrees
        Lref = Lref.cache_operator(np.asfortranarray(f(v)))
        wr = np.array(f(w0), order="F", copy=True)
        Lref.mul5(wr, np.asfortranarray(f(v)), 0.7, 0.3)
        check(sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy(), wr, dtype, f"{name} vs restated reference")


def test_update_coefficients(sb):
    """test/matrix.jl:223-265, test/basic.jl:405-426: time-dependent scalars and diagonals."""
    rng = np.random.default_rng(7)
    N, K = 16, 4
    A, d = rnd(rng, N, N), rnd(rng, N)
    v = rnd(rng, N, K)
    dv = dev(sb, v)
    lam = sb.ScalarOperator(0.0, update_func=lambda old, u, p, t: p * t)

    def upd(diag, u, p, t):
        diag.t.copy_(sb.DeviceArray.from_numpy(d * t).t)

    D = sb.DiagonalOperator(np.zeros(N), update_func=upd)
    L = sb.cache_operator(lam * sb.MatrixOperator(A) + D, dv)
    w = sb.DeviceArray.zeros((N, K), np.float64)
    for (p, t) in ((2.0, 3.0), (0.5, -1.0)):
        got = L(w, dv, None, p, t).numpy()
        check(got, (p * t * A + np.diag(d * t)) @ v, np.float64, f"p={p} t={t}")


def test_adapt_rehomes_every_array(sb):
    """test/adapt.jl:19-30,73-192: a storage-mapping function reaches every array of every operator type, the
    structure is preserved, and the adapted operator computes the same thing."""
    rng = np.random.default_rng(11)
    N, K = 12, 5
    A, d, Dk = rnd(rng, N, N), rnd(rng, N), rnd(rng, N, K)
    seen = []

    def to(a):
        seen.append(a.shape)
        return sb.DeviceArray(a.t.clone(), a.shape)

    L = (sb.DiagonalOperator(d) * sb.MatrixOperator(A).adjoint() + 2.0 * sb.MatrixOperator(A) + sb.IdentityOperator(N)
         + sb.DiagonalOperator(Dk) + sb.NullOperator(N))
    L2 = sb.adapt(L, to=to)
    assert sorted(seen) == sorted([(N,), (N, N), (N, N), (N, K)])
    assert type(L2) is type(L) and [type(o) for o in L2.ops] == [type(o) for o in L.ops]
    v = dev(sb, rnd(rng, N, K))
    check((sb.cache_operator(L2, v) * v).numpy(), (sb.cache_operator(L, v) * v).numpy(), np.float64, "adapted tree")
    seen.clear()
    T = sb.kron(sb.MatrixOperator(A), sb.kron(sb.IdentityOperator(3), sb.DiagonalOperator(d)))
    T2 = sb.adapt(T, to=to)
    assert sorted(seen) == sorted([(N, N), (N,)]) and T2.size() == T.size()
    vt = dev(sb, rnd(rng, T.size()[1], 3))
    check((sb.cache_operator(T2, vt) * vt).numpy(), (sb.cache_operator(T, vt) * vt).numpy(), np.float64, "adapted TPO")
    assert not sb.adapt(sb.cache_operator(sb.kron(A, A), dev(sb, rnd(rng, N * N, 2)))).iscached()   # caches are dropped


# ---- configuration shapes (BASELINE.json configs) ---------------------------------------------------------
def test_config1_full_size_vs_c_oracle(sb, oracle_c):
    """configs[0]: 100x100 (x) 100x100 Float64, V 10^4 x 100 -- full size, against the C restatement of both
    reference algorithms (stdlib permute path and LoopVectorization path)."""
    rng0, rng1, rng2 = (np.random.default_rng(s) for s in (0, 1, 2))
    A, B = rnd(rng0, 100, 100), rnd(rng0, 100, 100)
    v = rnd(rng1, 10 ** 4, 100)
    w0 = rnd(rng2, 10 ** 4, 100)
    dv = dev(sb, v)
    L = sb.cache_operator(sb.kron(A, B), dv)
    got3 = sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv).numpy()
    got5 = sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy()
    for lv in (True, False):
        check(got3, oracle_c.tpo2_mul(np.zeros_like(w0), A, B, v, lv=lv), np.float64, f"config1 3-arg lv={lv}")
        check(got5, oracle_c.tpo2_mul(w0.copy(order="F"), A, B, v, 0.7, 0.3, lv=lv), np.float64, f"config1 5-arg lv={lv}")


def test_config2_shapes_reduced_batch(sb):
    """configs[1] factors (512x512 (x) 512x512 Float64) on an 8-column batch against the mode-product oracle."""
    rng0, rng1, rng2 = (np.random.default_rng(s) for s in (0, 1, 2))
    A, B = rnd(rng0, 512, 512), rnd(rng0, 512, 512)
    k = 8
    v, w0 = rnd(rng1, 512 * 512, k), rnd(rng2, 512 * 512, k)
    dv = dev(sb, v)
    L = sb.cache_operator(sb.kron(A, B), dv)
    check(sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv).numpy(), R.kron_apply([A, B], v), np.float64, "config2 3-arg")
    check(sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy(), R.kron_apply([A, B], v, 0.7, 0.3, w0), np.float64, "config2 5-arg")


def test_config2_full_size_properties(sb):
    """configs[1] at full size (262144 x 256): size-independent properties -- sampled columns against the
    oracle, linearity in V, and the β == 0 no-read rule."""
    import torch
    rng0 = np.random.default_rng(0)
    A, B = rnd(rng0, 512, 512), rnd(rng0, 512, 512)
    n, k = 512 * 512, 256
    g = torch.Generator(device="cuda").manual_seed(1)
    V = sb.DeviceArray(torch.rand(n * k, dtype=torch.float64, device="cuda", generator=g), (n, k))
    L = sb.cache_operator(sb.kron(A, B), V)
    W = sb.DeviceArray.full((n, k), float("nan"), np.float64)
    sb.mul_(W, L, V)
    assert torch.isfinite(W.t).all()
    for j in (0, 101, 255):
        vj = V.cols(j, j + 1).numpy()
        check(W.cols(j, j + 1).numpy(), R.kron_apply([A, B], vj), np.float64, f"column {j}")
    # linearity: L(2V) == 2 L(V), and 5-arg with (α, β) = (2, -1) on W gives W again
    W2 = W.copy()
    sb.mul_(W2, L, V, 2.0, -1.0)
    err = (W2.t - W.t).norm() / W.t.norm()
    assert float(err) <= 1e-12, f"alpha/beta identity violated: {float(err):.3e}"


def test_config3_shapes_reduced_batch(sb):
    """configs[2]: 128x128 (x) 128x128 (x) 128x128 Float32 on a 2-column batch (full size needs 8.6 GB per
    array on the CPU side; the oracle is evaluated in Float64)."""
    rng0, rng1 = np.random.default_rng(0), np.random.default_rng(1)
    A, B, Cm = (rnd(rng0, 128, 128, dtype=np.float32) for _ in range(3))
    k = 2
    v = rnd(rng1, 128 ** 3, k, dtype=np.float32)
    dv = dev(sb, v)
    L = sb.cache_operator(sb.TensorProductOperator(A, B, Cm), dv)
    got = (L * dv).numpy()
    check(got, R.kron_apply([A, B, Cm], v), np.float32, "config3 3-arg")


def test_config3_full_size_properties(sb):
    """configs[2] at full size (2097152 x 512 Float32, 4.3 GB per array, tcgen05 path): sampled columns against
    the Float64 oracle, and the alpha/beta identity W <- 2 L V - W == W."""
    import torch
    rng0 = np.random.default_rng(0)
    F = [rnd(rng0, 128, 128, dtype=np.float32) for _ in range(3)]
    n, k = 128 ** 3, 512
    g = torch.Generator(device="cuda").manual_seed(1)
    V = sb.DeviceArray(torch.rand(n * k, dtype=torch.float32, device="cuda", generator=g), (n, k))
    L = sb.cache_operator(sb.TensorProductOperator(*F), V)
    W = sb.DeviceArray.full((n, k), float("nan"), np.float32)
    sb.mul_(W, L, V)
    assert torch.isfinite(W.t).all()
    for j in (0, 511):
        check(W.cols(j, j + 1).numpy(), R.kron_apply(F, V.cols(j, j + 1).numpy()), np.float32, f"column {j}")
    W2 = W.copy()
    sb.mul_(W2, L, V, 2.0, -1.0)
    err = float((W2.t - W.t).norm() / W.t.norm())
    assert err <= 1e-5, f"alpha/beta identity violated: {err:.3e}"
    del W2, W, V, L
    torch.cuda.empty_cache()


def test_config4_shapes_reduced(sb):
    """configs[3]: α(D*M' + 2M + I)v + βw, N = 2^12 (dense part), and the all-diagonal variant at N = 2^16."""
    rng = np.random.default_rng(4)
    N, K = 4096, 64
    Mm, d = rnd(rng, N, N), rnd(rng, N)
    v, w0 = rnd(rng, N, K), rnd(rng, N, K)
    dv = dev(sb, v)
    Mop = sb.MatrixOperator(Mm)
    L = sb.cache_operator(sb.DiagonalOperator(d) * Mop.adjoint() + 2.0 * Mop + sb.IdentityOperator(N), dv)
    want = 0.7 * (d[:, None] * (Mm.T @ v) + 2 * (Mm @ v) + v) + 0.3 * w0
    check(sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy(), want, np.float64, "config4 dense")
    N, K = 2 ** 16, 32
    d1, d2, Dk = rnd(rng, N), rnd(rng, N), rnd(rng, N, K)
    v, w0 = rnd(rng, N, K), rnd(rng, N, K)
    dv = dev(sb, v)
    L = sb.cache_operator(sb.DiagonalOperator(d1) * sb.DiagonalOperator(Dk) + 2.0 * sb.DiagonalOperator(d2)
                          + sb.IdentityOperator(N), dv)
    want = 0.7 * ((d1[:, None] * Dk + 2 * d2[:, None] + 1) * v) + 0.3 * w0
    check(sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy(), want, np.float64, "config4 elementwise")


def test_config5_laplacian_reduced(sb):
    """configs[4]: kron(Dxx, I) + kron(I, Dyy) on a 512x512 grid, 2 columns."""
    rng = np.random.default_rng(5)
    n = 512
    Dxx, Dyy = rnd(rng, n, n), rnd(rng, n, n)
    v, w0 = rnd(rng, n * n, 2), rnd(rng, n * n, 2)
    dv = dev(sb, v)
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(Dxx), sb.IdentityOperator(n))
                          + sb.kron(sb.IdentityOperator(n), sb.MatrixOperator(Dyy)), dv)
    want = np.empty_like(v)
    for j in range(2):
        Vj = v[:, j].reshape((n, n), order="F")
        want[:, j] = (Dyy @ Vj + Vj @ Dxx.T).reshape(-1, order="F")
    check(sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv).numpy(), want, np.float64, "laplacian 3-arg")
    check(sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy(), 0.7 * want + 0.3 * w0, np.float64, "laplacian 5-arg")
    Ls = sb.cache_operator(sb.kronsum(Dxx, Dyy), dv)
    check(sb.mul_(dev(sb, w0), Ls, dv, 0.7, 0.3).numpy(), 0.7 * want + 0.3 * w0, np.float64, "kronsum")


# ---- solver loop around the matvec -------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_col_dot_and_axpby(sb, dtype):
    rng = np.random.default_rng(21)
    for (n, k) in ((1000, 7), (4096, 3), (37, 5), (262144, 2)):
        X, Y, W0 = rnd(rng, n, k, dtype=dtype), rnd(rng, n, k, dtype=dtype), rnd(rng, n, k, dtype=dtype)
        an, ad, bn, bd = (rnd(rng, k, dtype=dtype) + 0.5 for _ in range(4))
        dX, dY = dev(sb, X), dev(sb, Y)
        f = lambda a: a.astype(np.float64)
        check(sb.krylov.col_dot(dX, dY).numpy(), (f(X) * f(Y)).sum(0), dtype, f"col_dot {n}x{k}")
        dW, dot = dev(sb, W0), sb.DeviceArray.full((k,), float("nan"), dtype)
        sb.krylov.col_axpby(dX, dW, sa=-0.7, an=dev(sb, an), ad=dev(sb, ad), sb=0.3, bn=dev(sb, bn), bd=dev(sb, bd), dot=dot)
        want = (-0.7 * f(an) / f(ad)) * f(X) + (0.3 * f(bn) / f(bd)) * f(W0)
        check(dW.numpy(), want, dtype, f"col_axpby {n}x{k}")
        check(dot.numpy(), (want * want).sum(0), dtype, f"fused dot {n}x{k}")
        dW = dev(sb, np.full_like(W0, np.nan))                      # sb == 0 never reads W
        sb.krylov.col_axpby(dX, dW, sa=2.0, sb=0.0)
        check(dW.numpy(), 2.0 * f(X), dtype, "col_axpby beta=0")


def test_batched_cg_device_resident(sb):
    """Batched CG on an SPD Kronecker sum (the 2-D Laplacian-like operator of config 5), every column an
    independent system: graph-captured iterations match a NumPy CG step for step and converge."""
    rng = np.random.default_rng(22)
    n1, n2, k = 24, 16, 5
    Q1, Q2 = rnd(rng, n1, n1), rnd(rng, n2, n2)
    A, B = Q1 @ Q1.T / n1 + np.eye(n1), Q2 @ Q2.T / n2 + np.eye(n2)          # SPD factors
    dense = np.kron(A, np.eye(n2)) + np.kron(np.eye(n1), B)
    Bm = rnd(rng, n1 * n2, k)
    iters = 16
    # NumPy reference, column by column
    Xr = np.zeros_like(Bm)
    for j in range(k):
        x = np.zeros(n1 * n2); r = Bm[:, j] - dense @ x; p = r.copy(); rs = r @ r
        for _ in range(iters):
            Ap = dense @ p; al = rs / (p @ Ap); x += al * p; r -= al * Ap; rn = r @ r; p = r + (rn / rs) * p; rs = rn
        Xr[:, j] = x
    for use_graph in (False, True):
        X = sb.DeviceArray.zeros(Bm.shape, np.float64)
        dB = dev(sb, Bm)
        L = sb.cache_operator(sb.kronsum(np.asfortranarray(A), np.asfortranarray(B)), dB)
        X, rs, _ = sb.krylov.cg(L, dB, X, iters, use_graph=use_graph, iters_per_graph=8)
        assert R.rel_err(X.numpy(), Xr) < 1e-9, f"graph={use_graph}"
        assert R.rel_err(dense @ X.numpy(), Bm) < 1e-6
        assert (rs.numpy() < 1e-10).all()


# ---- banded factors: finite-difference flavour of config 5 ------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_banded_factors(sb, dtype):
    rng = np.random.default_rng(31)

    def banded(n, b):
        A = rnd(rng, n, n, dtype=dtype)
        i, j = np.indices((n, n))
        A[np.abs(i - j) > b] = 0
        return A

    f = lambda a: np.asarray(a, dtype=np.float64)
    for (mo, bo, mi, bi, k) in ((9, 1, 13, 1, 5), (64, 2, 37, 1, 3), (5, 4, 6, 3, 2), (300, 1, 257, 1, 2), (40, 1, 300, 1, 3),
                                (600, 1, 520, 1, 2)):
        A, B = banded(mo, bo), banded(mi, bi)
        opA, opB = sb.MatrixOperator(sb.BandedMatrix.from_dense(A, bo)), sb.MatrixOperator(sb.BandedMatrix.from_dense(B, bi))
        v, w0 = rnd(rng, mo * mi, k, dtype=dtype), rnd(rng, mo * mi, k, dtype=dtype)
        dv = dev(sb, v)
        # banded leaf on its own
        vb = rnd(rng, mi, k, dtype=dtype)
        check((opB * dev(sb, vb)).numpy(), f(B) @ f(vb), dtype, "banded MatrixOperator")
        # Kronecker product with banded / dense / identity factors mixed (expected values through the
        # mode-product oracle: the dense Kronecker matrix of the largest case would be 47 GB)
        D = rnd(rng, mo, mo, dtype=dtype)
        for ops, mats in (((opA, opB), [A, B]), ((sb.MatrixOperator(D), opB), [D, B]),
                          ((opA, sb.IdentityOperator(mi)), [A, np.eye(mi)])):
            L = sb.cache_operator(sb.TensorProductOperator(*ops), dv)
            want = R.kron_apply([f(m) for m in mats], f(v))
            check(sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv).numpy(), want, dtype, "banded kron 3-arg")
            check(sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy(), 0.7 * want + 0.3 * f(w0), dtype, "banded kron 5-arg")
        # Kronecker sum = the finite-difference Laplacian: one fused pass
        S = sb.cache_operator(sb.kronsum(opA, opB), dv)
        assert S._banded
        want = R.kron_apply([f(A), np.eye(mi)], f(v)) + R.kron_apply([np.eye(mo), f(B)], f(v))
        check(sb.mul_(dev(sb, np.full_like(w0, np.nan)), S, dv).numpy(), want, dtype, "banded kronsum 3-arg")
        check(sb.mul_(dev(sb, w0), S, dv, 0.7, 0.3).numpy(), 0.7 * want + 0.3 * f(w0), dtype, "banded kronsum 5-arg")
        # the same operator spelled kron(A, I) + kron(I, B) (config 5's spelling) is recognised as that Kronecker sum
        for Ladd in (sb.kron(opA, sb.IdentityOperator(mi)) + sb.kron(sb.IdentityOperator(mo), opB),
                     sb.kron(sb.IdentityOperator(mo), opB) + sb.kron(opA, sb.IdentityOperator(mi))):
            Ladd = sb.cache_operator(Ladd, dv)
            assert Ladd._kronsum is not None and sb.iscached(Ladd)
            check(sb.mul_(dev(sb, w0), Ladd, dv, 0.7, 0.3).numpy(), 0.7 * want + 0.3 * f(w0), dtype, "added-TPO spelling")


# ---- end-to-end host entry --------------------------------------------------------------------------------
@pytest.mark.parametrize("beta", [None, 0.3])
def test_kron_mul_host_matches_device_path(sb, beta):
    import torch
    rng = np.random.default_rng(6)
    A, B = rnd(rng, 64, 48), rnd(rng, 32, 40)
    k = 37
    v = rnd(rng, 48 * 40, k)
    w0 = rnd(rng, 64 * 32, k)
    Vh = torch.from_numpy(v.reshape(-1, order="F").copy()).pin_memory()
    Wh = torch.from_numpy(w0.reshape(-1, order="F").copy()).pin_memory()
    Ah, Bh = np.asfortranarray(A), np.asfortranarray(B)
    h = sb.handle(0)
    ptrs = (C.c_void_p * 2)(Ah.ctypes.data, Bh.ctypes.data)
    dims = (C.c_int64 * 4)(64, 48, 32, 40)
    lds = (C.c_int64 * 2)(64, 32)
    kinds = (C.c_int32 * 2)(0, 0)
    a = C.byref(C.c_double(0.7))
    b = C.byref(C.c_double(beta)) if beta is not None else None
    rc = h.lib.scimlop_kron_mul_host(h.ptr, 0, 0, 2, ptrs, dims, lds, kinds, C.c_void_p(Vh.data_ptr()), 48 * 40,
                                     C.c_void_p(Wh.data_ptr()), 64 * 32, k, a, b, 5)
    h.check(rc, "kron_mul_host")
    got = Wh.numpy().reshape((64 * 32, k), order="F")
    want = 0.7 * (np.kron(A, B) @ v) + (beta * w0 if beta else 0.0)
    check(got, want, np.float64, "kron_mul_host")


# ---- error behaviour at the boundary ------------------------------------------------------------------------
def test_error_statuses(sb):
    import torch
    from scimlop_b200 import _lib
    h = sb.handle(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    A = dev(sb, np.eye(4))
    v = dev(sb, np.ones((16, 2)))
    w = dev(sb, np.ones((16, 2)))
    ptrs = (C.c_void_p * 2)(A.ptr, A.ptr)
    dims = (C.c_int64 * 4)(4, 4, 4, 4)
    lds = (C.c_int64 * 2)(4, 4)
    kinds = (C.c_int32 * 2)(0, 0)
    # workspace too small == operator not cached
    rc = h.lib.scimlop_kron_mul(h.ptr, 0, 0, 2, ptrs, dims, lds, kinds, C.c_void_p(v.ptr), 16, C.c_void_p(w.ptr), 16, 2,
                                None, None, None, 0, st)
    assert rc == _lib.ERR_WORKSPACE
    with pytest.raises(_lib.ScimlopError, match="workspace"):
        h.check(rc, "kron_mul")
    # null pointer, bad shape, bad dtype, bad precision, bad handle
    assert h.lib.scimlop_kron_mul(h.ptr, 0, 0, 2, ptrs, dims, lds, kinds, None, 16, C.c_void_p(w.ptr), 16, 2, None, None,
                                  None, 0, st) == _lib.ERR_NULL_POINTER
    assert h.lib.scimlop_kron_mul(h.ptr, 0, 0, 2, ptrs, dims, lds, kinds, C.c_void_p(v.ptr), 17, C.c_void_p(w.ptr), 16, 2,
                                  None, None, None, 0, st) == _lib.ERR_BAD_SHAPE
    assert h.lib.scimlop_matmul(h.ptr, 7, 0, 0, 4, 4, 1, C.c_void_p(A.ptr), 4, C.c_void_p(v.ptr), 4, C.c_void_p(w.ptr), 4,
                                None, None, st) == _lib.ERR_UNSUPPORTED
    assert h.lib.scimlop_matmul(h.ptr, 0, 2, 0, 4, 4, 1, C.c_void_p(A.ptr), 4, C.c_void_p(v.ptr), 4, C.c_void_p(w.ptr), 4,
                                None, None, st) == _lib.ERR_UNSUPPORTED
    assert h.lib.scimlop_matmul(h.ptr, 0, 0, 0, 4, 4, 1, C.c_void_p(A.ptr), 3, C.c_void_p(v.ptr), 4, C.c_void_p(w.ptr), 4,
                                None, None, st) == _lib.ERR_BAD_SHAPE
    assert h.lib.scimlop_matmul(None, 0, 0, 0, 4, 4, 1, C.c_void_p(A.ptr), 4, C.c_void_p(v.ptr), 4, C.c_void_p(w.ptr), 4,
                                None, None, st) == _lib.ERR_BAD_HANDLE
    # host-side protocol errors mirror the reference's @assert's
    with pytest.raises(AssertionError):
        sb.mul_(w, sb.MatrixOperator(np.eye(4)), v)             # dimension mismatch
    with pytest.raises(TypeError):
        sb.mul_(w, sb.MatrixOperator(np.eye(16, dtype=np.float32)), v)  # mixed eltypes are rejected
    # empty inputs are fine
    e = sb.DeviceArray.zeros((16, 0), np.float64)
    L = sb.cache_operator(sb.kron(np.eye(4), np.eye(4)), e)
    assert sb.mul_(sb.DeviceArray.zeros((16, 0), np.float64), L, e).shape == (16, 0)


# ---- few-column products (gemm_skinny.cuh): vector / thin-matrix mul! of a MatrixOperator ---------------------
SKINNY_CASES = [
    # M (rows of op(A)), N (reduction), k
    (256, 64, 1), (512, 256, 1), (300, 1000, 1), (4096, 4096, 1), (4100, 4104, 2), (1000, 77 * 2, 3), (2048, 3000, 5),
    (777 * 2, 1234, 8), (640, 5000, 9), (1026, 2050, 15), (64, 65536, 1), (48, 40000, 4), (16, 70000, 7),
    (10000, 128, 1), (20000, 64, 6), (2048, 2048, 16), (1500, 3000, 33), (520, 4000, 48),
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("trans", [False, True])
@pytest.mark.parametrize("case", SKINNY_CASES)
def test_matrix_operator_few_columns(sb, case, trans, dtype):
    """`mul!(w, L::MatrixOperator, v)` with a vector or a thin matrix (test/matrix.jl:72-80 uses K = 12 columns and
    vectors): the TMA-streamed, cluster-split path, 3-arg into NaN-filled w and 5-arg."""
    M, N, k = case
    rng = np.random.default_rng(M * 31 + N * 7 + k)
    A = rnd(rng, *((N, M) if trans else (M, N)), dtype=dtype) - dtype(0.5)
    v = rnd(rng, N, k, dtype=dtype) - dtype(0.5)
    w0 = rnd(rng, M, k, dtype=dtype)
    L = sb.MatrixOperator(A, trans=trans)
    opA = (A.T if trans else A).astype(np.float64)
    want = opA @ v.astype(np.float64)
    # summation error of a length-N Float32 dot product grows like sqrt(N) eps; the bound is the north_star's at N <= 4096
    tol = TOL[np.dtype(dtype)] * (max(1.0, np.sqrt(N / 4096.0)) if dtype == np.float32 else 1.0)
    w = sb.DeviceArray.full((M, k), float("nan"), dtype)
    sb.mul_(w, L, dev(sb, v))
    got = w.numpy()
    assert np.isfinite(got).all()
    # cancellation-free error measure: entries are sums of +-0.25-sized terms, compare against the norm of |A||v|
    scale = np.linalg.norm(np.abs(opA) @ np.abs(v.astype(np.float64)))
    assert np.linalg.norm(got - want) <= tol * scale, f"3-arg err {np.linalg.norm(got - want) / scale:.3e}"
    w = dev(sb, w0)
    sb.mul_(w, L, dev(sb, v), 0.7, 0.3)
    want5 = 0.7 * want + 0.3 * w0.astype(np.float64)
    assert np.linalg.norm(w.numpy() - want5) <= tol * (scale + np.linalg.norm(w0)), "5-arg"
    if k == 1:   # a true Vector input (ndim == 1)
        wv = sb.DeviceArray.full((M,), float("nan"), dtype)
        sb.mul_(wv, L, dev(sb, v[:, 0].copy()))
        assert np.linalg.norm(wv.numpy() - want[:, 0]) <= tol * scale


# ---- general Float32 GEMM on tcgen05 (streamed pre-split operand): large MatrixOperator products ---------------
TF32_STREAM_CASES = [
    # M, N, K, transA, transB
    (512, 256, 512, 0, 0), (512, 256, 512, 1, 0), (512, 256, 512, 0, 1), (512, 256, 512, 1, 1),
    (1000, 200, 1000, 0, 0), (300, 260, 700, 0, 0), (300, 260, 700, 1, 0), (2048, 16, 2048, 0, 0),
    (132, 1032, 380, 0, 1), (4096, 1024, 4096, 0, 0), (4096, 1024, 4096, 1, 0),
    # the FIRST operand is the small one (it is the one that gets pre-split; output contiguous along c)
    (512, 4096, 512, 0, 0), (512, 4096, 512, 1, 0), (512, 4096, 512, 0, 1), (512, 4096, 512, 1, 1),
    (200, 3000, 260, 0, 0), (16, 5000, 700, 1, 0),
]


@pytest.mark.parametrize("case", TF32_STREAM_CASES)
@pytest.mark.parametrize("precision,tol", [(0, 1e-5), (3, 1e-5), (2, 3e-3)])
def test_gemm_f32_streamed_tensor_core(sb, case, precision, tol):
    """`mul!(w, L::MatrixOperator{Float32}, v)` with a big matrix and many columns (config 4's dense part in
    Float32): arbitrary M, N, K on the tcgen05 3xTF32 kernel with the pre-split second operand streamed."""
    import torch
    M, N, K, ta, tb = case
    rng = np.random.default_rng(M + 3 * N + 7 * K + ta + 2 * tb)
    A = rnd(rng, *((K, M) if ta else (M, K)), dtype=np.float32) - np.float32(0.5)
    B = rnd(rng, *((N, K) if tb else (K, N)), dtype=np.float32) - np.float32(0.5)
    C0 = rnd(rng, M, N, dtype=np.float32)
    opA = (A.T if ta else A).astype(np.float64)
    opB = (B.T if tb else B).astype(np.float64)
    prod = opA @ opB
    scale = np.linalg.norm(np.abs(opA) @ np.abs(opB))      # cancellation-free reference magnitude
    h = sb.handle(0)
    for (a, b) in ((0.7, 0.3), (None, None)):
        dA, dB = dev(sb, A), dev(sb, B)
        dC = dev(sb, C0 if b else np.full_like(C0, np.nan))
        pa = C.byref(C.c_float(a)) if a is not None else None
        pb = C.byref(C.c_float(b)) if b is not None else None
        rc = h.lib.scimlop_gemm_strided_batched(
            h.ptr, 1, precision, ta, tb, M, N, K, pa, C.c_void_p(dA.ptr), A.shape[0], 0, C.c_void_p(dB.ptr), B.shape[0], 0,
            pb, C.c_void_p(dC.ptr), M, 0, 1, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        h.check(rc, "gemm f32 stream")
        got = dC.numpy().astype(np.float64)
        want = (1.0 if a is None else a) * prod + (b * C0 if b else 0.0)
        assert np.isfinite(got).all()
        err = np.linalg.norm(got - want) / (scale + (np.linalg.norm(C0) if b else 0.0))
        assert err <= tol, f"{case} precision={precision} alpha={a}: err {err:.3e}"


@pytest.mark.parametrize("factor", [256, 512, 200])
def test_tpo_float32_large_factors(sb, factor):
    """Float32 Kronecker product with factors beyond the resident-factor limit (128): both mode products run on the
    streamed tcgen05 kernel (stage 1 with the factor pre-split, stage 2 batched with the shared factor)."""
    rng = np.random.default_rng(factor)
    n, k = factor, 24
    A, B = rnd(rng, n, n, dtype=np.float32) - np.float32(0.5), rnd(rng, n, n, dtype=np.float32) - np.float32(0.5)
    v, w0 = rnd(rng, n * n, k, dtype=np.float32), rnd(rng, n * n, k, dtype=np.float32)
    dv = dev(sb, v)
    L = sb.cache_operator(sb.kron(A, B), dv)
    want = R.kron_apply([A.astype(np.float64), B.astype(np.float64)], v.astype(np.float64))
    scale = np.linalg.norm(R.kron_apply([np.abs(A).astype(np.float64), np.abs(B).astype(np.float64)], np.abs(v).astype(np.float64)))
    got = sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv).numpy()
    assert np.isfinite(got).all()
    assert np.linalg.norm(got - want) <= 1e-5 * scale
    got5 = sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy()
    assert np.linalg.norm(got5 - (0.7 * want + 0.3 * w0)) <= 1e-5 * (scale + np.linalg.norm(w0))


@pytest.mark.parametrize("K", [2048, 16384])
def test_gemm_f32_long_contraction_same_sign_data(sb, K):
    """Tensor-core FP32 accumulation rounds toward zero: on all-positive data an unchunked accumulator drifts by
    2e-8 * K (1.6e-4 at K = 8192).  The streamed kernel sums 128-k chunks outside the tensor core and must hold
    the Float32 tolerance at any contraction length."""
    rng = np.random.default_rng(K)
    M, N = 384, 256
    A, B = rnd(rng, M, K, dtype=np.float32), rnd(rng, K, N, dtype=np.float32)     # uniform [0, 1): every product positive
    L = sb.MatrixOperator(A)
    w = sb.DeviceArray.full((M, N), float("nan"), np.float32)
    sb.mul_(w, L, dev(sb, B))
    want = A.astype(np.float64) @ B.astype(np.float64)
    err = R.rel_err(w.numpy(), want)
    assert err <= 5e-6, f"K={K}: rel err {err:.3e}"


@pytest.mark.parametrize("N,K", [(1024, 64), (1536, 200)])
def test_config4_tree_float32_on_tensor_cores(sb, N, K):
    """configs[3] in Float32: α(D*M' + Dk*M + 2M + I)v + βw with the dense products on the streamed tcgen05 kernel and
    the vector / batched diagonal factors applied in its epilogue."""
    rng = np.random.default_rng(N + K)
    f32 = np.float32
    Mm = rnd(rng, N, N, dtype=f32) - f32(0.5)
    d, Dk = rnd(rng, N, dtype=f32), rnd(rng, N, K, dtype=f32)
    v, w0 = rnd(rng, N, K, dtype=f32) - f32(0.5), rnd(rng, N, K, dtype=f32)
    dv = dev(sb, v)
    Mop = sb.MatrixOperator(Mm)
    L = sb.cache_operator(sb.DiagonalOperator(d) * Mop.adjoint() + sb.DiagonalOperator(Dk) * Mop + 2.0 * Mop
                          + sb.IdentityOperator(N), dv)
    M64, v64 = Mm.astype(np.float64), v.astype(np.float64)
    want = 0.7 * (d[:, None] * (M64.T @ v64) + Dk * (M64 @ v64) + 2 * (M64 @ v64) + v64) + 0.3 * w0
    scale = np.linalg.norm(np.abs(M64) @ np.abs(v64)) * 4 + np.linalg.norm(w0)
    got = sb.mul_(dev(sb, w0), L, dv, 0.7, 0.3).numpy()
    assert np.isfinite(got).all()
    assert np.linalg.norm(got - want) <= 1e-5 * scale
    want3 = d[:, None] * (M64.T @ v64) + Dk * (M64 @ v64) + 2 * (M64 @ v64) + v64
    got3 = sb.mul_(dev(sb, np.full_like(w0, np.nan)), L, dv).numpy()
    assert np.linalg.norm(got3 - want3) <= 1e-5 * scale


@pytest.mark.parametrize("dtype,M,N,K,pad", [(np.float64, 300, 5, 1000, 4), (np.float64, 4096, 1, 4096, 2), (np.float32, 512, 256, 512, 4),
                                             (np.float32, 1000, 7, 2000, 8), (np.float64, 256, 64, 256, 6)])
@pytest.mark.parametrize("ta", [0, 1])
def test_gemm_padded_leading_dimensions(sb, dtype, M, N, K, pad, ta):
    """Sub-matrices of larger allocations (lda / ldb / ldc > rows) through every fast path: the skinny TMA kernel, the
    streamed tcgen05 kernel and the split-K DMMA kernel address operands by leading dimension only."""
    import torch
    rng = np.random.default_rng(M + N + K + ta)
    ar, ac = (K, M) if ta else (M, K)
    Abig = rnd(rng, ar + pad, ac, dtype=dtype) - dtype(0.5)
    Bbig = rnd(rng, K + pad, N, dtype=dtype) - dtype(0.5)
    Cbig0 = rnd(rng, M + pad, N, dtype=dtype)
    A, B = Abig[:ar], Bbig[:K]
    opA = (A.T if ta else A).astype(np.float64)
    want = 0.7 * (opA @ B.astype(np.float64)) + 0.3 * Cbig0[:M]
    scale = np.linalg.norm(np.abs(opA) @ np.abs(B.astype(np.float64))) + np.linalg.norm(Cbig0[:M])
    h = sb.handle(0)
    dA, dB, dC = dev(sb, Abig), dev(sb, Bbig), dev(sb, Cbig0)
    ct = C.c_double if dtype == np.float64 else C.c_float
    rc = h.lib.scimlop_gemm_strided_batched(
        h.ptr, 0 if dtype == np.float64 else 1, 0, ta, 0, M, N, K, C.byref(ct(0.7)), C.c_void_p(dA.ptr), ar + pad, 0,
        C.c_void_p(dB.ptr), K + pad, 0, C.byref(ct(0.3)), C.c_void_p(dC.ptr), M + pad, 0, 1,
        C.c_void_p(torch.cuda.current_stream().cuda_stream))
    h.check(rc, "gemm padded")
    got = dC.numpy()
    assert np.linalg.norm(got[:M] - want) <= TOL[np.dtype(dtype)] * scale
    assert np.array_equal(got[M:], Cbig0[M:]), "rows below the sub-matrix must not be touched"



def test_captured_mul_replays_with_updated_inputs(sb):
    """A captured mul! (CUDA graph) reads the CURRENT contents of v and of the operator's arrays at every replay."""
    import torch
    rng = np.random.default_rng(77)
    A, B = rnd(rng, 12, 12), rnd(rng, 9, 9)
    v1, v2 = rnd(rng, 108, 3), rnd(rng, 108, 3)
    dv = dev(sb, v1)
    w = sb.DeviceArray.full((108, 3), float("nan"), np.float64)
    opA = sb.MatrixOperator(A)
    L = sb.cache_operator(sb.kron(opA, B), dv)
    cm = sb.krylov.CapturedMul(L, w, dv)
    cm.replay(); torch.cuda.synchronize()
    check(w.numpy(), np.kron(A, B) @ v1, np.float64, "captured mul!")
    dv.t.copy_(dev(sb, v2).t)                                  # new input, same buffer
    opA.A.t.mul_(2.0)                                          # update_coefficients!-style in-place change of a factor
    cm.replay(); torch.cuda.synchronize()
    check(w.numpy(), np.kron(2 * A, B) @ v2, np.float64, "captured mul! after updates")


# ---- the no-allocation contract of the Float32 streamed path (test/Downstream/alloccheck.jl:14,100-122) ---------
def test_float32_large_operator_mul_is_one_launch_without_allocation(sb):
    """A cached large Float32 MatrixOperator: `cache_operator` prepares the hi / lo image once (scimlop_split_attach);
    every `mul!` after that is exactly ONE kernel launch and no device memory is allocated or freed."""
    import torch
    rng = np.random.default_rng(5)
    M, K, N = 640, 768, 512
    A = rnd(rng, M, K, dtype=np.float32) - np.float32(0.5)
    v = rnd(rng, K, N, dtype=np.float32)
    dv = dev(sb, v)
    h = sb.handle(0)
    Lop = sb.MatrixOperator(A, update_func=lambda Adev, u, p, t: Adev.t.mul_(float(t)))
    L = sb.cache_operator(Lop, dv)
    w = sb.DeviceArray.full((M, N), float("nan"), np.float32)
    sb.mul_(w, L, dv)                      # warm: lazy module loading etc.
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    n0 = h.launch_count()
    for _ in range(5):
        sb.mul_(w, L, dv)
    torch.cuda.synchronize()
    assert h.launch_count() - n0 == 5, "a cached Float32 mul! must be one launch (no per-call split)"
    free1, _ = torch.cuda.mem_get_info()
    assert free1 == free0, "mul! on a cached operator must not allocate device memory"
    want = A.astype(np.float64) @ v.astype(np.float64)
    scale = np.linalg.norm(np.abs(A).astype(np.float64) @ np.abs(v).astype(np.float64))
    assert np.linalg.norm(w.numpy() - want) <= 1e-5 * scale
    # update_coefficients! re-splits the changed matrix (one launch), the next mul! sees it
    n0 = h.launch_count()
    L.update_coefficients_(None, None, 3.0)
    assert h.launch_count() - n0 == 1
    sb.mul_(w, L, dv)
    assert np.linalg.norm(w.numpy() - 3.0 * want) <= 1e-5 * 3.0 * scale
    # the transposed operator shares the array and the image
    Lt = sb.cache_operator(sb.MatrixOperator(L.A, trans=True), dev(sb, rnd(rng, M, N, dtype=np.float32)))
    vt = rnd(rng, M, N, dtype=np.float32)
    wt = sb.DeviceArray.full((K, N), float("nan"), np.float32)
    n0 = h.launch_count()
    sb.mul_(wt, Lt, dev(sb, vt))
    assert h.launch_count() - n0 == 1
    assert np.linalg.norm(wt.numpy() - 3.0 * (A.astype(np.float64).T @ vt.astype(np.float64))) <= 1e-5 * 3.0 * np.linalg.norm(
        np.abs(A).astype(np.float64).T @ np.abs(vt).astype(np.float64))


def test_float32_kron_large_factors_launch_count(sb):
    """Float32 Kronecker product with 256-wide factors, cached: one launch per mode product, nothing else."""
    rng = np.random.default_rng(6)
    n, k = 256, 32
    A, B = rnd(rng, n, n, dtype=np.float32) - np.float32(0.5), rnd(rng, n, n, dtype=np.float32) - np.float32(0.5)
    v = rnd(rng, n * n, k, dtype=np.float32)
    dv = dev(sb, v)
    h = sb.handle(0)
    L = sb.cache_operator(sb.kron(A, B), dv)
    w = sb.DeviceArray.full((n * n, k), float("nan"), np.float32)
    sb.mul_(w, L, dv)
    n0 = h.launch_count()
    sb.mul_(w, L, dv)
    assert h.launch_count() - n0 == 2
    want = R.kron_apply([A.astype(np.float64), B.astype(np.float64)], v.astype(np.float64))
    scale = np.linalg.norm(R.kron_apply([np.abs(A).astype(np.float64), np.abs(B).astype(np.float64)], np.abs(v).astype(np.float64)))
    assert np.linalg.norm(w.numpy() - want) <= 1e-5 * scale


@pytest.mark.parametrize("attach", [False, True])
def test_float32_split_reads_only_the_logical_extent(sb, attach):
    """A sub-matrix view (lda > rows) that ends exactly at the end of its allocation: the hi / lo split may read
    ld * (cols - 1) + rows elements, not ld * cols (the padding behind the last column does not exist)."""
    import torch
    rng = np.random.default_rng(7)
    rows, cols, lda, N = 500, 384, 512, 640          # the view's last column stops 12 rows short of a full ld; A is the smaller operand
    Abig = rnd(rng, lda, cols, dtype=np.float32) - np.float32(0.5)
    logical = lda * (cols - 1) + rows
    flat = torch.from_numpy(np.ascontiguousarray(Abig.reshape(-1, order="F"))[:logical].copy()).cuda()   # exactly the logical extent
    v = rnd(rng, cols, N, dtype=np.float32)
    dv = dev(sb, v)
    h = sb.handle(0)
    w = sb.DeviceArray.full((rows, N), float("nan"), np.float32)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    img = None
    if attach:
        nb = C.c_size_t(0)
        h.lib.scimlop_split_bytes(rows, cols, C.byref(nb))
        img = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
        h.check(h.lib.scimlop_split_attach(h.ptr, C.c_void_p(flat.data_ptr()), lda, rows, cols, C.c_void_p(img.data_ptr()),
                                           nb.value, st), "attach")
    one = C.c_float(1.0)
    h.check(h.lib.scimlop_matmul(h.ptr, 1, 0, 0, rows, cols, N, C.c_void_p(flat.data_ptr()), lda, C.c_void_p(dv.ptr), cols,
                                 C.c_void_p(w.ptr), rows, C.byref(one), None, st), "matmul")
    torch.cuda.synchronize()
    if attach:
        h.lib.scimlop_split_detach(h.ptr, C.c_void_p(flat.data_ptr()))
    A = Abig[:rows]
    want = A.astype(np.float64) @ v.astype(np.float64)
    scale = np.linalg.norm(np.abs(A).astype(np.float64) @ np.abs(v).astype(np.float64))
    assert np.linalg.norm(w.numpy() - want) <= 1e-5 * scale


def test_captured_mul_sees_time_dependent_scalars(sb):
    """`update_coefficients!` of a time-dependent ScalarOperator between replays (src/scalar.jl:267-270): the kernels
    take λ·α by value, so the captured graph is re-captured when a scalar changed -- and only then."""
    import torch
    rng = np.random.default_rng(78)
    A, B = rnd(rng, 12, 12), rnd(rng, 9, 9)
    v, w0 = rnd(rng, 108, 3), rnd(rng, 108, 3)
    dv, w = dev(sb, v), dev(sb, w0)
    lam = sb.ScalarOperator(2.0, update_func=lambda old, u, p, t: 1.0 + t)
    L = sb.cache_operator(lam * sb.kron(A, B), dv)
    cm = sb.krylov.CapturedMul(L, w, dv, 0.5, 0.25)
    assert np.array_equal(w.numpy(), w0), "constructing a CapturedMul with beta != 0 must not modify w"
    dense = np.kron(A, B)
    cm.replay(); torch.cuda.synchronize()
    want = 0.5 * 2.0 * dense @ v + 0.25 * w0
    check(w.numpy(), want, np.float64, "captured 5-arg mul!")
    cm.replay(); torch.cuda.synchronize()
    want = 0.5 * 2.0 * dense @ v + 0.25 * want
    check(w.numpy(), want, np.float64, "second replay")
    assert cm.captures == 1
    L.update_coefficients_(None, None, 4.0)          # λ = 5 now
    cm.replay(); torch.cuda.synchronize()
    want = 0.5 * 5.0 * dense @ v + 0.25 * want
    check(w.numpy(), want, np.float64, "replay after update_coefficients! of the scalar")
    assert cm.captures == 2
    cm.replay(beta=0.0); torch.cuda.synchronize()     # new beta: 3-arg form, w never read
    check(w.numpy(), 0.5 * 5.0 * dense @ v, np.float64, "replay with a new beta")
    assert cm.captures == 3


def test_batched_cg_zero_rhs_column_stays_finite(sb):
    """An all-zero right-hand-side column: rs = pAp = 0, so the step lengths are 0/0 -- the vector kernels turn a zero
    denominator into a zero coefficient (that column stays at x = 0) instead of poisoning it with NaN."""
    rng = np.random.default_rng(23)
    n1, n2, k = 16, 12, 4
    Q1, Q2 = rnd(rng, n1, n1), rnd(rng, n2, n2)
    A, B = Q1 @ Q1.T / n1 + np.eye(n1), Q2 @ Q2.T / n2 + np.eye(n2)
    dense = np.kron(A, np.eye(n2)) + np.kron(np.eye(n1), B)
    Bm = rnd(rng, n1 * n2, k)
    Bm[:, 2] = 0.0
    for use_graph in (False, True):
        X = sb.DeviceArray.zeros(Bm.shape, np.float64)
        dB = dev(sb, Bm)
        L = sb.cache_operator(sb.kronsum(np.asfortranarray(A), np.asfortranarray(B)), dB)
        X, rs, _ = sb.krylov.cg(L, dB, X, 40, use_graph=use_graph, iters_per_graph=8)
        x = X.numpy()
        assert np.isfinite(x).all() and np.isfinite(rs.numpy()).all(), f"graph={use_graph}"
        assert np.array_equal(x[:, 2], np.zeros(n1 * n2))
        assert R.rel_err(dense @ x, Bm) < 1e-8


def test_dmma_tail_wave_runs_split_k_and_option_restores_one_launch(sb):
    """A Float64 GEMM with a few whole waves plus a partial one (the strong-scaled configs[1] share of one of 8 GPUs:
    512 x 512 (x) 512 x 512 on 32 columns = 512 tiles per stage on 148 SMs) runs the partial wave as a second,
    split-K launch: 2 launches per stage, same result within 1e-12; SCIMLOP_OPT_TAIL_SPLIT = 0 restores one launch per
    stage and the single summation order (bit-identical to the fused all-gather entries, bench.py's check)."""
    import torch
    rng = np.random.default_rng(11)
    A, B = rnd(rng, 512, 512), rnd(rng, 512, 512)
    n, k = 512 * 512, 32
    g = torch.Generator(device="cuda").manual_seed(5)
    V = sb.DeviceArray(torch.rand(n * k, dtype=torch.float64, device="cuda", generator=g), (n, k))
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)), V)
    h = sb.handle(0)
    W1 = sb.DeviceArray.full((n, k), float("nan"), np.float64)
    sb.mul_(W1, L, V)
    n0 = h.launch_count()
    sb.mul_(W1, L, V)
    split_launches = h.launch_count() - n0
    h.set_option(sb.OPT_TAIL_SPLIT, 0)
    try:
        W0 = sb.DeviceArray.full((n, k), float("nan"), np.float64)
        n0 = h.launch_count()
        sb.mul_(W0, L, V)
        assert h.launch_count() - n0 == 2
    finally:
        h.set_option(sb.OPT_TAIL_SPLIT, 1)
    assert split_launches == 4, f"expected main + tail launch per stage, got {split_launches}"
    assert torch.isfinite(W1.t).all()
    err = float((W1.t - W0.t).norm() / W0.t.norm())
    assert err <= 1e-14, f"split-K tail vs single launch: {err:.3e}"
    for j in (0, 31):
        check(W1.cols(j, j + 1).numpy(), R.kron_apply([A, B], V.cols(j, j + 1).numpy()), np.float64, f"column {j}")
    # 5-arg through the tail launch as well
    W2 = W1.copy()
    sb.mul_(W2, L, V, 2.0, -1.0)
    assert float((W2.t - W1.t).norm() / W1.t.norm()) <= 1e-12


@pytest.mark.parametrize("k,chunks,launches", [(32, 0, 4), (64, 0, 4), (64, 4, 8), (32, 4, 8), (8, 0, 2)])
def test_fused_all_gather_column_pipeline_is_bit_identical(sb, k, chunks, launches):
    """scimlop_kron_mul_peers (include/scimlop_b200.h) on ONE device: the "peer" copy of W is a second buffer of the same
    GPU, so the column-chunk pipeline (SCIMLOP_OPT_PEER_PIPELINE: last stage of chunk i beside the first stage of chunk
    i + 1, two streams, per-stage grid limits) runs without a second rank.  Both copies must equal the plain mul_ with
    one summation order (SCIMLOP_OPT_TAIL_SPLIT = 0) bit for bit, pipelined (2 chunks, or the 4 that
    SCIMLOP_OPT_PEER_CHUNKS asks for -- 128-tile chunks then stay off the few-tile split-K launch; none at 8 columns)
    and not, repeatedly; the launch count shows which schedule ran."""
    import types
    import torch
    from scimlop_b200 import distributed as sd
    rng = np.random.default_rng(23)
    A, B = rnd(rng, 512, 512), rnd(rng, 512, 512)
    n = 512 * 512
    g = torch.Generator(device="cuda").manual_seed(77)
    V = sb.DeviceArray(torch.rand(n * k, dtype=torch.float64, device="cuda", generator=g), (n, k))
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)), V)
    h = sb.handle(0)
    h.set_option(sb.OPT_TAIL_SPLIT, 0)
    try:
        want = sb.DeviceArray.zeros((n, k), np.float64)
        sb.mul_(want, L, V)
    finally:
        h.set_option(sb.OPT_TAIL_SPLIT, 1)
    if k == 8:   # few tiles: the first stage of the plain path runs split-K; compare to tolerance only
        check(want.cols(0, 1).numpy(), R.kron_apply([A, B], V.cols(0, 1).numpy()), np.float64, "plain")
    # "rank 0 of 2": own slab = columns 0 .. k of a 2k-column W, peer = another buffer on this device
    W0 = sb.DeviceArray.zeros((n, 2 * k), np.float64)
    W1 = sb.DeviceArray.zeros((n, 2 * k), np.float64)
    win = types.SimpleNamespace(W=W0, rank=0, world=2, base=[W0.ptr, W1.ptr])
    for pipelined in (2, 0, 2):   # 2 = always (automatic would decline: one local "peer" is not ingress-bound)
        h.set_option(sb.OPT_PEER_PIPELINE, pipelined)
        h.set_option(sb.OPT_PEER_CHUNKS, chunks)
        try:
            for rep in range(3):
                W0.t.zero_(); W1.t.zero_()
                n0 = h.launch_count()
                sd.kron_mul_peers(h, L, V, win, k, barrier=False)
                torch.cuda.synchronize()
                assert h.launch_count() - n0 == (launches if pipelined else 2)
                for name, Wx in (("own", W0), ("peer", W1)):
                    got = Wx.cols(0, k).t
                    if k == 8:
                        err = float((got - want.t).norm() / want.t.norm())
                        assert err <= 1e-14, f"{name} copy, pipelined={pipelined}: {err:.3e}"
                    else:
                        assert torch.equal(got, want.t), \
                            f"{name} copy, pipelined={pipelined}, rep {rep}: max diff {float((got - want.t).abs().max()):.3e}"
                    assert float(Wx.cols(k, 2 * k).t.abs().max()) == 0.0, "stored outside the slab"
        finally:
            h.set_option(sb.OPT_PEER_PIPELINE, 1)
            h.set_option(sb.OPT_PEER_CHUNKS, 0)
    # alpha rides along; beta is not part of the entry.  Automatic mode: unchunked here
    n0 = h.launch_count()
    sd.kron_mul_peers(h, L, V, win, k, alpha=-0.5, barrier=False)
    torch.cuda.synchronize()
    assert h.launch_count() - n0 == 2
    assert float((W1.cols(0, k).t + 0.5 * want.t).norm() / want.t.norm()) <= 1e-13


def test_mixed_eltype_factors_are_promoted(sb):
    """`_promote_operator_eltype` (src/interface.jl:444, src/tensor.jl:99): a Float32 factor next to Float64 arrays is
    applied at Float64 (exactly the promoted values), in a Kronecker product and in an operator tree."""
    rng = np.random.default_rng(12)
    A32, B = rnd(rng, 6, 5, dtype=np.float32), rnd(rng, 7, 4)
    v = rnd(rng, 20, 3)
    dv = dev(sb, v)
    L = sb.TensorProductOperator(A32, B)
    assert L.dtype == np.float64
    L = sb.cache_operator(L, dv)
    check((L * dv).numpy(), np.kron(A32.astype(np.float64), B) @ v, np.float64, "promoted kron")
    M32, D = rnd(rng, 9, 9, dtype=np.float32), rnd(rng, 9, 9)
    T = sb.cache_operator(sb.MatrixOperator(M32) + 2.0 * sb.MatrixOperator(D), dev(sb, rnd(rng, 9, 2)))
    x = rnd(rng, 9, 2)
    check((T * dev(sb, x)).numpy(), (M32.astype(np.float64) + 2.0 * D) @ x, np.float64, "promoted tree")


def test_copy_shares_no_storage_and_stays_cached(sb):
    """`Base.copy(L)` (src/tensor.jl:210-215): ops and cache are copied; mutating the original leaves the copy alone."""
    rng = np.random.default_rng(13)
    A, B = rnd(rng, 5, 5), rnd(rng, 6, 6)
    v = rnd(rng, 30, 4)
    dv = dev(sb, v)
    L = sb.cache_operator(sb.TensorProductOperator(A, B), dv)
    Lc = sb.copy(L)
    assert sb.iscached(Lc)
    want = np.kron(A, B) @ v
    L.ops[0].A.t.zero_()                      # wreck the original's outer factor in place
    check((Lc * dv).numpy(), want, np.float64, "copy after the original changed")
    assert np.abs((L * dv).numpy()).max() == 0.0
    Tc = sb.copy(sb.cache_operator(sb.MatrixOperator(A) + sb.DiagonalOperator(rnd(rng, 5)), dev(sb, rnd(rng, 5, 2))))
    assert sb.iscached(Tc)


@pytest.mark.parametrize("case", [((100, 100), (100, 100), 100), ((9, 13), (17, 11), 5), ((128, 128), (128, 128), 3),
                                  ((8, 8), (8, 8), 1), ((33, 64), (120, 50), 250), ((100, 37), (64, 128), 7),
                                  # 13 and 14 strips of 8 in either dimension: the strips beyond the twelfth are cut among
                                  # warps 12..15 (kron2_f64_small.cuh, assign_strips); 15 strips are not
                                  ((97, 104), (104, 97), 9), ((112, 105), (109, 112), 4), ((105, 100), (100, 113), 6),
                                  ((120, 99), (101, 120), 3)])
def test_small_float64_kron_is_one_fused_launch(sb, case):
    """configs[0]-like problems (two dense Float64 factors <= 128, fewer than two waves of slabs): ONE launch of
    kron2_f64_small_kernel with the intermediate in shared memory; 3-arg (NaN-prefilled W), 5-arg, adjoint factors."""
    (ma, na), (mb, nb), k = case
    rng = np.random.default_rng(ma + nb + k)
    A, B = rnd(rng, ma, na) - 0.5, rnd(rng, mb, nb) - 0.5
    v = rnd(rng, na * nb, k) - 0.5
    dv = dev(sb, v)
    h = sb.handle(0)
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)), dv)
    w = sb.DeviceArray.full((ma * mb, k), float("nan"), np.float64)
    sb.mul_(w, L, dv)
    n0 = h.launch_count()
    sb.mul_(w, L, dv)
    assert h.launch_count() - n0 == 1
    want = R.kron_apply([A, B], v)
    check(w.numpy(), want, np.float64, f"{case} 3-arg")
    w0 = rnd(rng, ma * mb, k)
    check(sb.mul_(dev(sb, w0), L, dv, 0.7, -0.3).numpy(), 0.7 * want - 0.3 * w0, np.float64, f"{case} 5-arg")
    vt = rnd(rng, ma * mb, k) - 0.5
    dvt = dev(sb, vt)
    Lt = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)).adjoint(), dvt)
    check((Lt * dvt).numpy(), R.kron_apply([A.T.copy(), B.T.copy()], vt), np.float64, f"{case} adjoint")
