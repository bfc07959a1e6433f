"""Times every BASELINE.json config (plus the sub-variants SURVEY.md §8(d) asks for) on one B200 through the
host mirror / C ABI, device-resident, CUDA events, and prints one JSON line per config with the roofline
fraction that bounds it.  Not the headline bench (that is bench.py); this is the per-path scoreboard that
goes into profiles/.

    python tools/bench_configs.py [--only c2,c4elt] [--reps 10]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb  # noqa: E402

PEAKS = {"f64_dmma_tflops": 37.1, "f32_ffma_tflops": 72.3}
try:
    _mp = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    PEAKS["hbm_gbs"] = float(_mp["hbm_gbs"])
    PEAKS["hbm_source"] = "MEASURED_PEAKS.json"
except Exception:
    PEAKS["hbm_gbs"] = 6650.0
    PEAKS["hbm_source"] = "fallback (B200_PROFILING.md)"


def dev_rand(shape, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    n = int(np.prod(shape))
    t = torch.rand(n, dtype=torch.float64 if dtype == np.float64 else torch.float32, device="cuda", generator=g)
    return sb.DeviceArray(t, shape)


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    return float(np.min(times)), float(np.median(times))


def report(name, flops, nbytes, tmin, tmed, bound, extra=None):
    h = sb.handle()
    line = {"config": name, "ms_min": tmin, "ms_median": tmed, "tflops": flops / (tmed * 1e-3) / 1e12,
            "gbs": nbytes / (tmed * 1e-3) / 1e9, "bound": bound}
    if bound == "hbm":
        line["frac"] = line["gbs"] / PEAKS["hbm_gbs"]
        line["peak"] = f'{PEAKS["hbm_gbs"]} GB/s ({PEAKS["hbm_source"]})'
    elif bound == "f64":
        line["frac"] = line["tflops"] / PEAKS["f64_dmma_tflops"]
        line["peak"] = "37.1 TFLOP/s FP64 DMMA (profiles/r01_peaks_microbench.json)"
    else:
        line["frac"] = line["tflops"] / PEAKS["f32_ffma_tflops"]
        line["peak"] = "72.3 TFLOP/s FP32 FFMA (profiles/r01_peaks_microbench.json)"
    line["launches_total"] = h.launch_count()
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def c1(reps):
    A, B = dev_rand((100, 100), np.float64, 0), dev_rand((100, 100), np.float64, 10)
    V = dev_rand((10 ** 4, 100), np.float64, 1)
    W = dev_rand((10 ** 4, 100), np.float64, 2)
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)), V)
    for tag, fn in (("3arg", lambda: sb.mul_(W, L, V)), ("5arg", lambda: sb.mul_(W, L, V, 0.7, 0.3))):
        t = timeit(fn, reps)
        report(f"c1 100x100(x)100x100 f64 k=100 {tag}", 4.0e8, 1.616e7 + (8e6 if tag == "5arg" else 0), *t, "f64")
    # the same mul! as a captured CUDA graph (krylov.CapturedMul: what a solver loop replays): the eager number above is
    # bound by the host work between two 10-us launches, not by the kernels
    cm = sb.krylov.CapturedMul(L, W, V)
    t = timeit(cm.replay, reps)
    report("c1 100x100(x)100x100 f64 k=100 3arg, CUDA-graph replay", 4.0e8, 1.616e7, *t, "f64")
    # 20 replays inside ONE graph: per-matvec device time without any launch gap
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(20):
                sb.mul_(W, L, V)
    torch.cuda.current_stream().wait_stream(side)
    t = timeit(g.replay, reps)
    report("c1 100x100(x)100x100 f64 k=100 3arg, 20 matvecs in one graph (per matvec)", 4.0e8, 1.616e7, t[0] / 20, t[1] / 20, "f64")


def c2(reps):
    A, B = dev_rand((512, 512), np.float64, 0), dev_rand((512, 512), np.float64, 10)
    V = dev_rand((512 * 512, 256), np.float64, 1)
    W = dev_rand((512 * 512, 256), np.float64, 2)
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)), V)
    for tag, fn in (("3arg", lambda: sb.mul_(W, L, V)), ("5arg", lambda: sb.mul_(W, L, V, 0.7, 0.3))):
        t = timeit(fn, reps)
        report(f"c2 512x512(x)512x512 f64 k=256 {tag}", 1.374e11, 1.078e9 + (5.37e8 if tag == "5arg" else 0), *t, "f64")


def c3(reps, k=512):
    """configs[2].  HBM-bound: the roofline fraction is ALGORITHMIC bytes (V read once, W written once; + W read for
    the 5-arg form) over the measured copy bandwidth.  Passes actually made: 2 (modes C and B fused on chip by
    kron2_f16_kernel, then mode A), so the traffic floor of this schedule is 2x the algorithmic bytes."""
    F = [dev_rand((128, 128), np.float32, s) for s in (0, 10, 20)]
    n = 128 ** 3
    V = dev_rand((n, k), np.float32, 1)
    W = sb.DeviceArray.empty((n, k), np.float32)
    L = sb.cache_operator(sb.TensorProductOperator(*[sb.MatrixOperator(f) for f in F]), V)
    t = timeit(lambda: sb.mul_(W, L, V), max(2, reps // 3), warm=1)
    report(f"c3 128^3 three-factor f32 k={k} 3arg", 3 * 2 * 128.0 * n * k, 2.0 * n * k * 4, *t, "hbm",
           {"hbm_frac_of_2_pass_traffic": 2 * 2.0 * n * k * 4 / (t[1] * 1e-3) / 1e9 / PEAKS["hbm_gbs"]})
    W.t.fill_(0.5)
    t = timeit(lambda: sb.mul_(W, L, V, 0.7, 0.3), max(2, reps // 3), warm=1)
    report(f"c3 128^3 three-factor f32 k={k} 5arg", 3 * 2 * 128.0 * n * k, 3.0 * n * k * 4, *t, "hbm",
           {"hbm_frac_of_2_pass_traffic": (2 * 2.0 + 1.0) * n * k * 4 / (t[1] * 1e-3) / 1e9 / PEAKS["hbm_gbs"]})


def c3pair(reps, k=32768):
    """The fused two-mode kernel alone: 128x128 (x) 128x128 Float32 on a 16384 x k batch (one launch)."""
    F = [dev_rand((128, 128), np.float32, s) for s in (0, 10)]
    n = 128 ** 2
    V = dev_rand((n, k), np.float32, 1)
    W = sb.DeviceArray.empty((n, k), np.float32)
    L = sb.cache_operator(sb.TensorProductOperator(*[sb.MatrixOperator(f) for f in F]), V)
    t = timeit(lambda: sb.mul_(W, L, V), max(2, reps // 3), warm=1)
    report(f"c3pair 128^2 two-factor f32 k={k} 3arg (kron2_f16_kernel)", 2 * 2 * 128.0 * n * k, 2.0 * n * k * 4, *t, "hbm",
           {"us_per_slab_per_sm": t[1] * 1e3 / (k / 148.0)})
    W.t.fill_(0.5)
    t = timeit(lambda: sb.mul_(W, L, V, 0.7, 0.3), max(2, reps // 3), warm=1)
    report(f"c3pair 128^2 two-factor f32 k={k} 5arg (kron2_f16_kernel)", 2 * 2 * 128.0 * n * k, 3.0 * n * k * 4, *t, "hbm")


def c4elt(reps):
    N, K = 2 ** 16, 1024
    d1, d2 = dev_rand((N,), np.float64, 3), dev_rand((N,), np.float64, 4)
    Dk = dev_rand((N, K), np.float64, 5)
    V, W = dev_rand((N, K), np.float64, 1), dev_rand((N, K), np.float64, 2)
    L = sb.cache_operator(sb.DiagonalOperator(d1) * sb.DiagonalOperator(Dk) + 2.0 * sb.DiagonalOperator(d2)
                          + sb.IdentityOperator(N), V)
    nb = N * K * 8
    t = timeit(lambda: sb.mul_(W, L, V, 0.7, 0.3), reps)
    report("c4-elt alpha*(D1*Dk + 2*D2 + I)*v + beta*w f64 N=65536 K=1024", 5.0 * N * K, 4.0 * nb, *t, "hbm")
    t = timeit(lambda: sb.mul_(W, L, V), reps)
    report("c4-elt 3-arg (no read of w)", 4.0 * N * K, 3.0 * nb, *t, "hbm")
    Lb = sb.DiagonalOperator(Dk)
    t = timeit(lambda: sb.mul_(W, Lb, V, 0.7, 0.3), reps)
    report("BatchedDiagonalOperator 5-arg", 3.0 * N * K, 4.0 * nb, *t, "hbm")
    Ld = sb.DiagonalOperator(d1)
    t = timeit(lambda: sb.mul_(W, Ld, V), reps)
    report("DiagonalOperator 3-arg", 1.0 * N * K, 2.0 * nb, *t, "hbm")
    Li = sb.IdentityOperator(N)
    t = timeit(lambda: sb.mul_(W, Li, V, 0.7, 0.3), reps)
    report("IdentityOperator 5-arg (axpby)", 2.0 * N * K, 3.0 * nb, *t, "hbm")


def c4dense(reps, N=65536, K=1024):   # configs[3] at its named size: the matrix is 34 GB
    Mm = dev_rand((N, N), np.float64, 6)
    d = dev_rand((N,), np.float64, 3)
    V, W = dev_rand((N, K), np.float64, 1), dev_rand((N, K), np.float64, 2)
    Mop = sb.MatrixOperator(Mm)
    L = sb.cache_operator(sb.DiagonalOperator(d) * Mop.adjoint() + 2.0 * Mop + sb.IdentityOperator(N), V)
    t = timeit(lambda: sb.mul_(W, L, V, 0.7, 0.3), max(2, reps // 4), warm=1)
    report(f"c4-dense alpha*(D*M' + 2M + I)*v + beta*w f64 N={N} K={K}", 2 * 2.0 * N * N * K,
           2.0 * N * N * 8 + 3.0 * N * K * 8, *t, "f64")


def c5(reps, k=8):
    n = 4096
    Dxx, Dyy = dev_rand((n, n), np.float64, 7), dev_rand((n, n), np.float64, 8)
    V, W = dev_rand((n * n, k), np.float64, 1), dev_rand((n * n, k), np.float64, 2)
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(Dxx), sb.IdentityOperator(n))
                          + sb.kron(sb.IdentityOperator(n), sb.MatrixOperator(Dyy)), V)
    t = timeit(lambda: sb.mul_(W, L, V), max(2, reps // 4), warm=1)
    report(f"c5 kron(Dxx,I)+kron(I,Dyy) 4096^2 grid f64 k={k} (AddedOperator of two TPOs)", 2 * 2.0 * n ** 3 * k,
           2.0 * n * n * k * 8 + 2.0 * n * n * 8, *t, "f64")
    Ls = sb.cache_operator(sb.kronsum(sb.MatrixOperator(Dxx), sb.MatrixOperator(Dyy)), V)
    t = timeit(lambda: sb.mul_(W, Ls, V), max(2, reps // 4), warm=1)
    report(f"c5 kronsum(Dxx,Dyy) k={k} (scimlop_kronsum_mul)", 2 * 2.0 * n ** 3 * k, 2.0 * n * n * k * 8 + 2.0 * n * n * 8,
           *t, "f64")
    # the Krylov/ODE loop: R matvecs captured in one CUDA graph
    R = 20
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        sb.mul_(W, Ls, V)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(R):
                sb.mul_(W, Ls, V)
    t = timeit(lambda: g.replay(), 3, warm=1)
    report(f"c5 kronsum k={k}: {R} matvecs in one CUDA graph (per matvec)", 2 * 2.0 * n ** 3 * k,
           2.0 * n * n * k * 8 + 2.0 * n * n * 8, t[0] / R, t[1] / R, "f64")


def krylov(reps, k=8):
    """Solver-loop kernels at config-5 size (16.7M x 8) and a graph-captured batched CG iteration on the
    4096^2-grid Kronecker sum (SPD factors)."""
    n = 4096
    N = n * n
    X, Y, W = dev_rand((N, k), np.float64, 1), dev_rand((N, k), np.float64, 2), dev_rand((N, k), np.float64, 3)
    out = sb.DeviceArray.zeros((k,), np.float64)
    nb = N * k * 8
    t = timeit(lambda: sb.krylov.col_dot(X, Y, out), reps)
    report(f"col_dot {N}x{k} f64", 2.0 * N * k, 2.0 * nb, *t, "hbm")
    t = timeit(lambda: sb.krylov.col_axpby(X, W, sa=0.5, an=out, ad=out, sb=1.0), reps)
    report(f"col_axpby (device coefficients) {N}x{k} f64", 2.0 * N * k, 3.0 * nb, *t, "hbm")
    t = timeit(lambda: sb.krylov.col_axpby(X, W, sa=0.5, sb=1.0, dot=out), reps)
    report(f"col_axpby + fused |w|^2 {N}x{k} f64", 4.0 * N * k, 3.0 * nb, *t, "hbm")
    g = torch.Generator(device="cuda").manual_seed(5)
    def spd():
        Q = torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g)
        M = Q @ Q.T / n + torch.eye(n, dtype=torch.float64, device="cuda")
        return sb.DeviceArray(M.T.contiguous().reshape(-1), (n, n))
    L = sb.kronsum(sb.MatrixOperator(spd()), sb.MatrixOperator(spd()))
    Bm = dev_rand((N, k), np.float64, 4)
    Xs = sb.DeviceArray.zeros((N, k), np.float64)
    _, _, st = sb.krylov.cg(L, Bm, Xs, 8, use_graph=True, iters_per_graph=8)      # builds the graph
    t = timeit(lambda: st.graph.replay(), 3, warm=1)
    flops_it = 2 * 2.0 * n ** 3 * k
    report(f"batched CG iteration (graph of 8) on kronsum 4096^2 grid k={k}: per iteration", flops_it,
           11.0 * nb + 3.0 * nb, t[0] / 8, t[1] / 8, "f64", {"note": "matvec = 2 DMMA launches; 4 vector kernels per iteration"})


def c5fd(reps, k=8):
    """Finite-difference flavour of config 5: tridiagonal Dxx, Dyy on the 4096^2 grid -- one HBM-bound pass."""
    n = 4096
    N = n * n
    def tri(seed):
        b = dev_rand((n, 3), np.float64, seed)
        return sb.MatrixOperator(sb.BandedMatrix(b, 1))
    V, W = dev_rand((N, k), np.float64, 1), dev_rand((N, k), np.float64, 2)
    L = sb.cache_operator(sb.kronsum(tri(7), tri(8)), V)
    nb = N * k * 8
    t = timeit(lambda: sb.mul_(W, L, V), reps)
    report(f"c5-fd kronsum(tridiag, tridiag) 4096^2 grid f64 k={k} 3-arg", 10.0 * N * k, 2.0 * nb, *t, "hbm")
    t = timeit(lambda: sb.mul_(W, L, V, 0.7, 0.3), reps)
    report(f"c5-fd kronsum(tridiag, tridiag) 4096^2 grid f64 k={k} 5-arg", 12.0 * N * k, 3.0 * nb, *t, "hbm")
    L2 = sb.cache_operator(sb.kron(tri(7), sb.IdentityOperator(n)) + sb.kron(sb.IdentityOperator(n), tri(8)), V)
    t = timeit(lambda: sb.mul_(W, L2, V), reps)
    report(f"c5-fd as AddedOperator of two banded TPOs (two passes) k={k}", 10.0 * N * k, 5.0 * nb, *t, "hbm")


def csr(reps, N=65536, K=1024, per_row=16):
    """Sparse MatrixOperator (CSR gather kernel): N x N with ~per_row nonzeros per row on an N x K Float64 batch.
    Algorithmic bytes: V read once + W written once + the CSR arrays once per 8-column tile (they are re-read per tile)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    cols = torch.randint(0, N, (N, per_row), device="cuda", generator=g, dtype=torch.int32)
    cols, _ = torch.sort(cols, dim=1)
    rowptr = torch.arange(0, N * per_row + 1, per_row, device="cuda", dtype=torch.int32)
    vals = torch.rand(N * per_row, device="cuda", generator=g, dtype=torch.float64)
    A = sb.SparseMatrixCSR(sb.DeviceArray(rowptr, (N + 1,)), sb.DeviceArray(cols.reshape(-1).contiguous(), (N * per_row,)),
                           sb.DeviceArray(vals, (N * per_row,)), (N, N))
    L = sb.MatrixOperator(A)
    V = dev_rand((N, K), np.float64, 1)
    W = sb.DeviceArray.empty((N, K), np.float64)
    nnz = N * per_row
    t = timeit(lambda: sb.mul_(W, L, V), reps)
    report(f"csr N={N} nnz/row={per_row} f64 k={K} 3arg, direct gather (uncached)", 2.0 * nnz * K, 2.0 * N * K * 8 + nnz * 12.0 * (K / 8), *t, "hbm",
           {"dense_bytes_only_gbs": 2.0 * N * K * 8 / (t[1] * 1e-3) / 1e9})
    L = sb.cache_operator(L, V)
    t = timeit(lambda: sb.mul_(W, L, V), reps)
    report(f"csr N={N} nnz/row={per_row} f64 k={K} 3arg, cached (row-contiguous staging of V)", 2.0 * nnz * K,
           4.0 * N * K * 8 + nnz * 12.0 * (K / 8), *t, "hbm", {"dense_bytes_only_gbs": 2.0 * N * K * 8 / (t[1] * 1e-3) / 1e9})


def spkron(reps, N=4096, per_row=16, nd=128, K=64):
    """A general sparse matrix as a Kronecker FACTOR (SCIMLOP_CSR): S (N x N, per_row nonzeros per row) next to a dense
    nd x nd factor, Float64, K columns -- S as the outer factor (gather over contiguous mode rows of nd entries) and as
    the inner one (the staged SpMM on nd * K columns).  Algorithmic bytes: input + output + ONE intermediate round trip."""
    g = torch.Generator(device="cuda").manual_seed(4)
    cols, _ = torch.sort(torch.randint(0, N, (N, per_row), device="cuda", generator=g, dtype=torch.int32), dim=1)
    rowptr = torch.arange(0, N * per_row + 1, per_row, device="cuda", dtype=torch.int32)
    vals = torch.rand(N * per_row, device="cuda", generator=g, dtype=torch.float64)
    S = sb.MatrixOperator(sb.SparseMatrixCSR(sb.DeviceArray(rowptr, (N + 1,)), sb.DeviceArray(cols.reshape(-1).contiguous(), (N * per_row,)),
                                             sb.DeviceArray(vals, (N * per_row,)), (N, N)))
    B = sb.MatrixOperator(dev_rand((nd, nd), np.float64, 5))
    n = N * nd
    V = dev_rand((n, K), np.float64, 1)
    W = sb.DeviceArray.empty((n, K), np.float64)
    flops = (2.0 * N * per_row * nd + 2.0 * nd * nd * N) * K
    for tag, ops in (("S (x) B (sparse outer factor)", (S, B)), ("B (x) S (sparse inner factor, staged)", (B, S))):
        L = sb.cache_operator(sb.TensorProductOperator(*ops), V)
        t = timeit(lambda: sb.mul_(W, L, V), reps)
        report(f"spkron {tag}: S {N}^2 {per_row} nnz/row, B {nd}^2 dense, f64 k={K}", flops, 4.0 * n * K * 8, *t, "hbm")


def c3bf16(reps, k=32768):
    """The opt-in BF16 single-pass mode of the fused two-mode kernel (SCIMLOP_PREC_BF16)."""
    F = [dev_rand((128, 128), np.float32, s) for s in (0, 10)]
    n = 128 ** 2
    V = dev_rand((n, k), np.float32, 1)
    W = sb.DeviceArray.empty((n, k), np.float32)
    L = sb.cache_operator(sb.TensorProductOperator(*[sb.MatrixOperator(f) for f in F]), V)
    L.precision = 4
    t = timeit(lambda: sb.mul_(W, L, V), max(2, reps // 3), warm=1)
    report(f"c3pair 128^2 two-factor f32 k={k} 3arg, SCIMLOP_PREC_BF16 (opt-in)", 2 * 2 * 128.0 * n * k, 2.0 * n * k * 4, *t, "hbm")


def c2c(reps, k=128):
    """Complex variant of config 2: 512 x 512 (x) 512 x 512 ComplexF64 on 262144 x 128 (same bytes as the Float64 config):
    4 planar real DMMA GEMMs per mode between a de-interleave and an interleave pass."""
    g = torch.Generator(device="cuda").manual_seed(2)
    def crand(shape):
        n = int(np.prod(shape))
        t = torch.complex(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5,
                          torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5)
        return sb.DeviceArray(t, shape)
    A, B = crand((512, 512)), crand((512, 512))
    V, W = crand((512 * 512, k)), crand((512 * 512, k))
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)), V)
    t = timeit(lambda: sb.mul_(W, L, V), reps)
    report(f"c2 complex: 512x512(x)512x512 ComplexF64 k={k} 3arg", 4 * 1.374e11 * k / 256, 2 * 512.0 * 512 * k * 16, *t, "f64")


ALL = {"c2c": c2c, "c3bf16": c3bf16, "c5fd": c5fd, "krylov": krylov, "c1": c1, "c2": c2, "c3": c3, "c3pair": c3pair, "csr": csr, "spkron": spkron, "c4elt": c4elt, "c4dense": c4dense, "c5": c5}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=",".join(ALL))
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--c3k", type=int, default=512, help="columns for config 3 (smaller for ncu runs)")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    for name in a.only.split(","):
        if name == "c3":
            c3(a.reps, a.c3k)
        else:
            ALL[name](a.reps)
        torch.cuda.empty_cache()
