"""GPU parity of the fused two-mode Float32 Kronecker kernel (csrc/kron2_f16.cuh) against the Float64 oracle:
the `TensorProductOperator` mul! of src/tensor.jl:543-610 / 750-834 for small dense factors (configs[2] of BASELINE.json).
Every call goes through the C ABI (scimlop_kron_mul); tolerance 1e-5 norm-wise, the north_star's Float32 bound."""
import numpy as np
import pytest

from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def sb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    import scimlop_b200 as sb
    sb.handle(0)
    return sb


def nrm(rng, *shape):
    return np.asfortranarray(rng.standard_normal(shape).astype(np.float32))


def dev(sb, a):
    return sb.DeviceArray.from_numpy(a)


def apply(sb, factors, v, alpha=None, beta=None, w0=None):
    dv = dev(sb, v)
    L = sb.cache_operator(sb.TensorProductOperator(*factors), dv)
    rows = int(np.prod([f.shape[0] for f in factors]))
    if alpha is None:
        w = sb.DeviceArray.full((rows, v.shape[1]), float("nan"), np.float32)
        sb.mul_(w, L, dv)
    else:
        w = dev(sb, w0)
        sb.mul_(w, L, dv, alpha, beta)
    return w.numpy()


def err_vs_oracle(got, factors, v, alpha=None, beta=None, w0=None):
    want = R.kron_apply([f.astype(np.float64) for f in factors], v.astype(np.float64))
    if alpha is not None:
        want = alpha * want + beta * w0.astype(np.float64)
    assert np.isfinite(got).all(), "non-finite output"
    return R.rel_err(got, want)


def test_identity_factors_reproduce_the_input(sb):
    """F1 = F2 = I: data movement, slab scales and the FP16 split round trip (|err| <= 2^-22 per entry)."""
    rng = np.random.default_rng(0)
    I = np.asfortranarray(np.eye(128, dtype=np.float32))
    v = nrm(rng, 128 * 128, 5)
    got = apply(sb, [I, I], v)
    assert np.abs(got - v).max() <= 2.0 ** -21 * np.abs(v).max()


@pytest.mark.parametrize("which", ["inner", "outer"])
def test_one_dense_factor_at_a_time(sb, which):
    """I (x) C exercises stage 1 alone, B (x) I stage 2 alone."""
    rng = np.random.default_rng(1)
    I = np.asfortranarray(np.eye(128, dtype=np.float32))
    F = nrm(rng, 128, 128)
    v = nrm(rng, 128 * 128, 3)
    # IdentityOperator factors are skipped by the planner; a dense identity matrix keeps both modes on the fused kernel
    factors = [I, F] if which == "inner" else [F, I]
    e = err_vs_oracle(apply(sb, factors, v), factors, v)
    assert e <= TOL, f"{which}: rel err {e:.3e}"


@pytest.mark.parametrize("k", [1, 3, 7])
def test_two_factor_128(sb, k):
    rng = np.random.default_rng(2)
    A, B = nrm(rng, 128, 128), nrm(rng, 128, 128)
    v = nrm(rng, 128 * 128, k)
    e = err_vs_oracle(apply(sb, [A, B], v), [A, B], v)
    assert e <= TOL, f"3-arg rel err {e:.3e}"
    w0 = nrm(rng, 128 * 128, k)
    e = err_vs_oracle(apply(sb, [A, B], v, 0.7, -1.3, w0), [A, B], v, 0.7, -1.3, w0)
    assert e <= TOL, f"5-arg rel err {e:.3e}"


def test_two_factor_launch_count_and_no_read_of_w(sb):
    rng = np.random.default_rng(3)
    A, B = nrm(rng, 128, 128), nrm(rng, 128, 128)
    v = nrm(rng, 128 * 128, 4)
    dv = dev(sb, v)
    h = sb.handle(0)
    L = sb.cache_operator(sb.TensorProductOperator(A, B), dv)
    w = sb.DeviceArray.full((128 * 128, 4), float("nan"), np.float32)
    sb.mul_(w, L, dv)
    n0 = h.launch_count()
    sb.mul_(w, L, dv)
    assert h.launch_count() - n0 == 1, "two small Float32 factors are ONE fused launch"
    assert np.isfinite(w.numpy()).all(), "beta == 0 must not read W (NaN-prefilled)"


@pytest.mark.parametrize("dims", [((96, 128), (128, 100)), ((128, 96), (100, 128)), ((100, 100), (100, 100)),
                                  ((128, 104), (112, 120))])
def test_rectangular_and_ragged_factors(sb, dims):
    """Extents below 128 ride on TMA zero fill / clipped stores; (m, n) of the outer and the inner factor."""
    rng = np.random.default_rng(4)
    (ma, na), (mb, nb) = dims
    A, B = nrm(rng, ma, na), nrm(rng, mb, nb)
    v = nrm(rng, na * nb, 5)
    e = err_vs_oracle(apply(sb, [A, B], v), [A, B], v)
    assert e <= TOL, f"{dims}: rel err {e:.3e}"
    w0 = nrm(rng, ma * mb, 5)
    e = err_vs_oracle(apply(sb, [A, B], v, -2.0, 0.5, w0), [A, B], v, -2.0, 0.5, w0)
    assert e <= TOL, f"{dims} 5-arg: rel err {e:.3e}"


def test_adjoint_factors(sb):
    rng = np.random.default_rng(5)
    A, B = nrm(rng, 128, 112), nrm(rng, 104, 128)
    vt = nrm(rng, 128 * 104, 3)
    dvt = dev(sb, vt)
    Lt = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)).adjoint(), dvt)
    got = (Lt * dvt).numpy()
    want = R.kron_apply([A.T.astype(np.float64), B.T.astype(np.float64)], vt.astype(np.float64))
    assert R.rel_err(got, want) <= TOL


def test_slab_scales_cover_the_float32_range(sb):
    """Every slab is scaled by its own power of two: columns 1e-30 ... 1e+30 apart, an all-zero column, and a slab
    whose entries span 12 decades all hold the tolerance COLUMN BY COLUMN."""
    rng = np.random.default_rng(6)
    A, B = nrm(rng, 128, 128), nrm(rng, 128, 128)
    k = 8
    v = nrm(rng, 128 * 128, k)
    scales = np.array([1e-30, 1e-12, 1.0, 1e9, 1e30, 0.0, 1.0, 3e-38], dtype=np.float32)
    v *= scales[None, :]
    v[:, 6] *= np.float32(10.0) ** rng.uniform(-6, 6, size=128 * 128).astype(np.float32)
    got = apply(sb, [A * np.float32(1e-20), B * np.float32(1e15)], v)
    want = R.kron_apply([A.astype(np.float64) * np.float64(np.float32(1e-20)), B.astype(np.float64) * np.float64(np.float32(1e15))],
                        v.astype(np.float64))
    # the factors are rounded to Float32 after scaling: rebuild them exactly as passed
    Af, Bf = (A * np.float32(1e-20)).astype(np.float64), (B * np.float32(1e15)).astype(np.float64)
    want = R.kron_apply([Af, Bf], v.astype(np.float64))
    assert np.isfinite(got).all()
    for j in range(k):
        if scales[j] == 0.0:
            assert (got[:, j] == 0).all()
            continue
        if scales[j] < 1e-37:
            continue   # results of the denormal column underflow in Float32 itself; finiteness is the check
        e = R.rel_err(got[:, j], want[:, j])
        assert e <= TOL, f"column {j} (scale {scales[j]:g}): rel err {e:.3e}"


def test_three_factors_fused_pair_then_outer_mode(sb):
    """configs[2] shape on a small batch: modes (C, B) fused, mode A by the resident-factor kernel: 2 launches."""
    rng = np.random.default_rng(7)
    F = [nrm(rng, 128, 128) for _ in range(3)]
    k = 2
    v = nrm(rng, 128 ** 3, k)
    dv = dev(sb, v)
    h = sb.handle(0)
    L = sb.cache_operator(sb.TensorProductOperator(*F), dv)
    w = sb.DeviceArray.full((128 ** 3, k), float("nan"), np.float32)
    sb.mul_(w, L, dv)
    n0 = h.launch_count()
    sb.mul_(w, L, dv)
    assert h.launch_count() - n0 == 2
    e = err_vs_oracle(w.numpy(), F, v)
    assert e <= TOL, f"rel err {e:.3e}"
    w0 = nrm(rng, 128 ** 3, k)
    dw = dev(sb, w0)
    sb.mul_(dw, L, dv, 1.5, 0.25)
    e = err_vs_oracle(dw.numpy(), F, v, 1.5, 0.25, w0)
    assert e <= TOL, f"5-arg rel err {e:.3e}"


@pytest.mark.parametrize("chunks", [2, 4])
def test_three_factor_two_pass_pipeline_is_bit_identical(sb, chunks):
    """SCIMLOP_OPT_MODE_PIPELINE (include/scimlop_b200.h): the fused two-mode pass and the third-mode pass of three small
    dense Float32 factors run as a pipeline over column chunks on two streams with an SM share each.  Columns are
    independent, so the result must equal the unchunked call bit for bit (3- and 5-argument forms, repeated), stay
    within 1e-5 of the Float64 oracle, and shapes the pipeline does not take (a column count the chunk count does not
    divide; 64-wide factors below the fused kernel's size) must run unchanged."""
    import torch
    rng = np.random.default_rng(17 + chunks)
    F = [nrm(rng, 128, 128) for _ in range(3)]
    k = 8
    v = nrm(rng, 128 ** 3, k)
    dv = dev(sb, v)
    L = sb.cache_operator(sb.TensorProductOperator(*F), dv)
    h = sb.handle(0)
    w_ref = sb.DeviceArray.full((128 ** 3, k), float("nan"), np.float32)
    sb.mul_(w_ref, L, dv)
    w0 = nrm(rng, 128 ** 3, k)
    w5_ref = dev(sb, w0)
    sb.mul_(w5_ref, L, dv, 0.5, -0.25)
    h.set_option(sb.OPT_MODE_PIPELINE, chunks)
    try:
        for rep in range(3):
            w = sb.DeviceArray.full((128 ** 3, k), float("nan"), np.float32)
            n0 = h.launch_count()
            sb.mul_(w, L, dv)
            assert h.launch_count() - n0 == 2 * chunks, "one fused launch + one third-mode launch per chunk"
            torch.cuda.synchronize()
            assert torch.equal(w.t, w_ref.t), f"{chunks} chunks, rep {rep}: differs from the unchunked call"
            w5 = dev(sb, w0)
            sb.mul_(w5, L, dv, 0.5, -0.25)
            assert torch.equal(w5.t, w5_ref.t)
        # not taken: k = 7 is not divisible; the call must still be the plain two launches and the same numbers
        v7 = np.asfortranarray(v[:, :7])
        dv7 = dev(sb, v7)
        L7 = sb.cache_operator(sb.TensorProductOperator(*F), dv7)
        w7 = sb.DeviceArray.full((128 ** 3, 7), float("nan"), np.float32)
        n0 = h.launch_count()
        sb.mul_(w7, L7, dv7)
        assert h.launch_count() - n0 == 2
        assert torch.equal(w7.t, w_ref.t[: 128 ** 3 * 7])
        # not taken: 64-wide factors
        G = [nrm(rng, 64, 64) for _ in range(3)]
        vg = nrm(rng, 64 ** 3, 8)
        got = apply(sb, G, vg)
        assert R.rel_err(got, R.kron_apply([g.astype(np.float64) for g in G], vg.astype(np.float64))) <= TOL
    finally:
        h.set_option(sb.OPT_MODE_PIPELINE, 0)
    for j in (0, k - 1):
        want = R.kron_apply([f.astype(np.float64) for f in F], v[:, j:j + 1].astype(np.float64))
        assert R.rel_err(w_ref.cols(j, j + 1).numpy(), want) <= TOL


def test_same_sign_data_accumulation(sb):
    """All-positive factors and data: the tensor core's truncating FP32 accumulation is at its worst."""
    rng = np.random.default_rng(8)
    A = np.asfortranarray(rng.random((128, 128)).astype(np.float32))
    B = np.asfortranarray(rng.random((128, 128)).astype(np.float32))
    v = np.asfortranarray(rng.random((128 * 128, 4)).astype(np.float32))
    e = err_vs_oracle(apply(sb, [A, B], v), [A, B], v)
    assert e <= TOL, f"rel err {e:.3e}"


def test_precision_modes_bypass_the_fused_kernel(sb):
    """SCIMLOP_PREC_FP32_EXACT keeps the FFMA path (bit-level Float32 semantics), launch count 2."""
    rng = np.random.default_rng(9)
    A, B = nrm(rng, 128, 128), nrm(rng, 128, 128)
    v = nrm(rng, 128 * 128, 2)
    dv = dev(sb, v)
    h = sb.handle(0)
    L = sb.cache_operator(sb.TensorProductOperator(A, B), dv)
    L.precision = 3   # SCIMLOP_PREC_FP32_EXACT (include/scimlop_b200.h)
    w = sb.DeviceArray.full((128 * 128, 2), float("nan"), np.float32)
    sb.mul_(w, L, dv)
    n0 = h.launch_count()
    sb.mul_(w, L, dv)
    assert h.launch_count() - n0 == 2
    assert err_vs_oracle(w.numpy(), [A, B], v) <= 1e-6


def test_bf16_opt_in_mode(sb):
    """SCIMLOP_PREC_BF16 (explicit opt-in): one tcgen05 pass on BF16-rounded operands, FP32 accumulation.  Error of the
    order of the BF16 unit roundoff (2^-9) per mode, far from the default mode's 1e-6 -- and finite / unscaled for data
    far outside the FP16 range (BF16 keeps the FP32 exponent)."""
    rng = np.random.default_rng(10)
    A, B = nrm(rng, 128, 128), nrm(rng, 128, 128)
    v = nrm(rng, 128 * 128, 3)
    v[:, 1] *= np.float32(1e20)
    dv = dev(sb, v)
    h = sb.handle(0)
    L = sb.cache_operator(sb.TensorProductOperator(A, B), dv)
    L.precision = 4   # SCIMLOP_PREC_BF16
    w = sb.DeviceArray.full((128 * 128, 3), float("nan"), np.float32)
    sb.mul_(w, L, dv)
    n0 = h.launch_count()
    sb.mul_(w, L, dv)
    assert h.launch_count() - n0 == 1
    got = w.numpy()
    want = R.kron_apply([A.astype(np.float64), B.astype(np.float64)], v.astype(np.float64))
    for j in range(3):
        e = R.rel_err(got[:, j], want[:, j])
        assert 1e-5 < e <= 1e-2, f"column {j}: rel err {e:.3e} is not BF16-like"
    w0 = nrm(rng, 128 * 128, 3)
    dw = dev(sb, w0)
    sb.mul_(dw, L, dv, 0.5, 2.0)
    e = R.rel_err(dw.numpy()[:, 0], 0.5 * want[:, 0] + 2.0 * w0[:, 0].astype(np.float64))
    assert e <= 1e-2
    # three factors: the fused pair runs BF16, the third mode the single-pass TF32 kernel
    F = [nrm(rng, 128, 128) for _ in range(3)]
    v3 = nrm(rng, 128 ** 3, 1)
    dv3 = dev(sb, v3)
    L3 = sb.cache_operator(sb.TensorProductOperator(*F), dv3)
    L3.precision = 4
    e = err_vs_oracle((L3 * dv3).numpy(), F, v3)
    assert e <= 1e-2, f"3-factor BF16 mode: {e:.3e}"


def test_bf16_mode_many_slabs_per_cta(sb):
    """The BF16 mode on a batch large enough for every CTA to run dozens of slabs (the splitter groups share the raw ring:
    they must stay within one slab of each other in this mode too)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(3)
    k = 148 * 24
    F = [sb.DeviceArray(torch.rand(128 * 128, dtype=torch.float32, device="cuda", generator=g) - 0.5, (128, 128)) for _ in range(2)]
    V = sb.DeviceArray(torch.rand(128 * 128 * k, dtype=torch.float32, device="cuda", generator=g) - 0.5, (128 * 128, k))
    L = sb.cache_operator(sb.TensorProductOperator(*[sb.MatrixOperator(f) for f in F]), V)
    L.precision = 4
    W = sb.DeviceArray.full((128 * 128, k), float("nan"), np.float32)
    for _ in range(3):
        sb.mul_(W, L, V)
    torch.cuda.synchronize()
    assert torch.isfinite(W.t).all()
    for j in (0, k // 2, k - 1):
        want = R.kron_apply([f.numpy().astype(np.float64) for f in F], V.cols(j, j + 1).numpy().astype(np.float64))
        assert R.rel_err(W.cols(j, j + 1).numpy(), want) <= 2e-2
    L.precision = 0
    sb.mul_(W, L, V)
    for j in (1, k - 2):
        want = R.kron_apply([f.numpy().astype(np.float64) for f in F], V.cols(j, j + 1).numpy().astype(np.float64))
        assert R.rel_err(W.cols(j, j + 1).numpy(), want) <= 1e-5
