#!/usr/bin/env python
"""bench.py -- headline benchmark of the kron(A,B)·V `mul!` path (BASELINE.json metric) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (configs[1] of BASELINE.json): TensorProductOperator(MatrixOperator(A), MatrixOperator(B)) with
512x512 Float64 factors, `cache_operator` once, then `mul!(w, L, v)` on a 262144 x 256 batch.  A "step" is
one `mul!` over the whole batch.  Synthetic inputs: uniform [0,1), NumPy default_rng seed 0 (factors, outer
first), 1 (V), 2 (W0) -- BASELINE.md §2.

Multi-GPU (one process per GPU, torchrun): columns are independent units, every rank owns a full
262144 x 256 slab (weak scaling, global batch = 256*N columns), factors replicated, NO collective on the
data path.  `value` stays that weak number; the same line also carries, under `config.strong`,
the STRONG-scaled named shape (256 columns in total, 256/N per rank through scimlop_shard_columns) -- compute only,
and with W replicated on every rank (compute -> ncclAllGather, and the fused compute + all-gather kernels over
peer stores / NVSwitch multicast, each bit-compared with the unfused result: a mismatch fails the run) -- and,
under `config.c5`, configs[4] (dense 4096^2 Kronecker sum, one column per GPU, R = 50 matvecs in one CUDA graph).

Prints ONE JSON line (rank 0).  `value` = whole-job TFLOP/s with V, W resident in HBM; `e2e` = the same metric
through the host-buffer C-ABI call (pinned host V/W, H2D and D2H inside the timed region); `roofline` = the
DMMA GEMM kernel against the measured FP64-DMMA issue peak; `cpu_baseline` = the CPU restatement of the
reference algorithm (oracle/) timed on this box's host cores.

`--impl reference` times the reference's own CPU algorithm (Julia is not installable here, so this is the
oracle port: OpenBLAS GEMM + materialised permutes, src/tensor.jl:568,778-788) on the SAME configuration: the
full 262144 x 256 batch per step, all host threads."""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

if "--impl" in sys.argv and "reference" in sys.argv:
    # the CPU arm uses every host thread it can, also under torchrun (which exports OMP_NUM_THREADS=1)
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MO = NO = MI = NI = 512          # outer A (mo x no), inner B (mi x ni)
KCOLS = 256                      # columns per GPU
FLOPS_PER_COL = 2.0 * MI * NI * NO + 2.0 * MO * NO * MI      # stage 1 + stage 2
METRIC = "kron(A,B)*V mul! throughput (512x512 (x) 512x512 Float64, 262144x256 batch per GPU)"
FP64_DMMA_PEAK_TFLOPS = 37.1     # measured on this pool: tools/peaks.cu -> profiles/r01_peaks_microbench.json
# dram__bytes_read.sum + dram__bytes_write.sum per launch of gemm_dmma_kernel on this workload, from the
# `ncu --set full` capture summarised in profiles/r01_ncu_dmma_v6_summary.txt (539.0 + 492.1 MB stage 1,
# 539.2 + 495.2 MB stage 2): the algorithmic 537 MB in + 537 MB out, i.e. no re-reads.
NCU_DRAM_BYTES_PER_LAUNCH = 0.5 * ((540.313856 + 493.162752) + (539.751424 + 495.018496)) * 1e6   # profiles/r02_ncu_dmma_summary.txt
REF_COLS = KCOLS                 # --impl reference / cpu_baseline (stdlib variant): the full batch, as benchmarks/tensor.jl:17-24 times it
LV_SAMPLE_COLS = 32              # cpu_baseline, LoopVectorization variant (single-thread nest): columns timed, unscaled
C5_N = 4096                      # configs[4]: Dxx, Dyy 4096 x 4096 dense Float64
C5_MATVECS = 50                  # matvecs captured in one CUDA graph (R >= 50)


def make_factors():
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.random((MO, NO)))
    B = np.asfortranarray(rng.random((MI, NI)))
    return A, B


# --------------------------------------------------------------------------------------------------
# CPU arm: restatement of the reference's algorithm (oracle/), all host threads
# --------------------------------------------------------------------------------------------------
def cpu_time_one(variant, A, B, v, reps):
    """min wall time (what @btime reports) of one cached mul!(w, L, v) on the CPU restatement."""
    from oracle import ref_numpy as R
    L = R.TensorProductOperator(A, B, lv=False).cache_operator(v)
    w = np.zeros((MO * MI, v.shape[1]), order="F")
    if variant == "stdlib":
        fn = lambda: L.mul3(w, v)
    else:
        from oracle import cwrap
        lib = cwrap.load()
        c1 = L.cache[0]

        def fn():
            # stage 1 through BLAS (src/tensor.jl:568), stage 2 through the single-thread SIMD nest
            # (ext/SciMLOperatorsLoopVectorizationExt.jl:17-23)
            np.matmul(B, v.reshape((NI, NO * v.shape[1]), order="F"), out=c1)
            lib.lv_outer_mul(w, A, c1, MI, MO, NO, v.shape[1])
    best = float("inf")
    fn()                                   # warm-up (page faults of the scratch, BLAS thread start-up)
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    A, B = make_factors()
    k = REF_COLS
    v = np.asfortranarray(np.random.default_rng(1).random((NO * NI, k)))
    from oracle import ref_numpy as R
    L = R.TensorProductOperator(A, B, lv=False).cache_operator(v)
    w = np.zeros((MO * MI, k), order="F")
    for _ in range(args.warmup):
        L.mul3(w, v)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        L.mul3(w, v)
    dt = (time.perf_counter() - t0) / args.steps
    tflops = FLOPS_PER_COL * k / dt / 1e12
    cores = os.cpu_count()
    sample = f"all {k} of {KCOLS} columns per step (the whole configs[1] batch; nothing scaled)"
    line = {
        "impl": "reference", "metric": METRIC, "value": tflops, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: TensorProductOperator(MatrixOperator(A), MatrixOperator(B)), A,B 512x512 "
                               "Float64, cache_operator + mul!(w,L,v) on a 262144x256 batch per GPU",
                   "implementation": "CPU restatement of the reference algorithm (OpenBLAS gemm + materialised "
                                     "permutedims, src/tensor.jl:568,778-788), all host threads -- Julia is not "
                                     "installable in this image", "sample": sample},
        "cpu_baseline": {"value": tflops, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tflops, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region (NVML)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.ok = [], set(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.ok:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# extra legs of the GPU arm (reported inside `config`; `value` stays the weak-scaled headline)
# --------------------------------------------------------------------------------------------------
def _timed_max(torch, dist, world, fn, reps, warm=3):
    """ms per call: CUDA events on the current stream, barrier on both sides, max over ranks."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def copy_floor_ms(torch, dist, world, Vh, Wh, Vd, Wd, reps=5):
    """Pure pinned H2D of V concurrent with D2H of W (two streams, no kernels): the PCIe / host-memory floor of the
    e2e step on this box at this rank count (all ranks copy at the same time)."""
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s_up):
            Vd.copy_(Vh, non_blocking=True)
        with torch.cuda.stream(s_dn):
            Wh.copy_(Wd, non_blocking=True)
    once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3


def strong_legs(torch, dist, sb, h, world, rank, A, B, ms_1gpu_shape):
    """configs[1] STRONG-scaled: 256 columns in total, columns [first, first+count) per rank.  Returns a dict for
    `config.strong`; raises SystemExit(3) if a fused all-gather variant is not bit-identical to the unfused result."""
    from scimlop_b200 import distributed as sd
    n = NO * NI
    first, kper = sd.shard_columns(KCOLS, world, rank)
    assert kper * world == KCOLS

    def slab_of(r):
        g = torch.Generator(device="cuda").manual_seed(4242 + r)
        return sb.DeviceArray(torch.rand(n * kper, dtype=torch.float64, device="cuda", generator=g), (n, kper))

    V = slab_of(rank)
    Wl = sb.DeviceArray.full((n, kper), float("nan"), np.float64)
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)), V)
    out = {"columns_total": KCOLS, "cols_per_rank": kper, "flops_total": FLOPS_PER_COL * KCOLS}
    t_comp = _timed_max(torch, dist, world, lambda: sb.mul_(Wl, L, V), reps=20)
    out["compute_only_ms"] = t_comp
    out["compute_only_tflops"] = FLOPS_PER_COL * KCOLS / (t_comp * 1e-3) / 1e12
    out["speedup_vs_1gpu"] = ms_1gpu_shape / t_comp if ms_1gpu_shape else None
    if world == 1:
        return out

    # ---- W replicated on every rank ----
    def bcast(b):
        o = [b]
        dist.broadcast_object_list(o, src=0)
        return o[0]

    def gather_obj(o):
        res = [None] * world
        dist.all_gather_object(res, o)
        return res

    sd.init_comm(h, world, rank, bcast)
    W = sb.DeviceArray.zeros((n, KCOLS), np.float64)
    mine = W.cols(first, first + kper)
    out["gathered_bytes_per_rank"] = (world - 1) * n * kper * 8
    out["replicated_nccl_ms"] = _timed_max(torch, dist, world, lambda: (sb.mul_(mine, L, V), sd.allgather_cols(h, W, kper)), reps=20)

    # the unfused result of EVERY slab, recomputed locally from that rank's (seeded) V: the bit-compare reference
    def check(Wfull, what):
        # same summation order as the fused kernels: one launch per stage (the split-K tail launch reorders the sums
        # of the last partial wave; include/scimlop_b200.h, SCIMLOP_OPT_TAIL_SPLIT)
        ref = sb.DeviceArray.empty((n, kper), np.float64)
        h.set_option(sb.OPT_TAIL_SPLIT, 0)
        try:
            for r in range(world):
                sb.mul_(ref, L, slab_of(r))
                if not torch.equal(Wfull.cols(r * kper, (r + 1) * kper).t, ref.t):
                    d = float((Wfull.cols(r * kper, (r + 1) * kper).t - ref.t).abs().max())
                    print(json.dumps({"error": f"{what}: slab {r} on rank {rank} differs from the unfused result", "max_abs_diff": d}),
                          file=sys.stderr, flush=True)
                    return False
        finally:
            h.set_option(sb.OPT_TAIL_SPLIT, 1)
        return True

    ok = True
    try:
        win = sd.PeerWindow(h, W, rank, world, gather_obj)
        out["replicated_fused_peer_ms"] = _timed_max(torch, dist, world, lambda: sd.kron_mul_peers(h, L, V, win, kper), reps=20)
        h.set_option(sb.OPT_PEER_PIPELINE, 0)      # both stages over all columns in turn (the round-1 schedule)
        out["replicated_fused_peer_unpipelined_ms"] = _timed_max(torch, dist, world, lambda: sd.kron_mul_peers(h, L, V, win, kper), reps=20)
        h.set_option(sb.OPT_PEER_PIPELINE, 1)
        W.t.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        sd.kron_mul_peers(h, L, V, win, kper)
        torch.cuda.synchronize()
        dist.barrier()
        good = check(W, "fused peer-store all-gather")
        out["replicated_fused_peer_bit_identical"] = good
        ok = ok and good
        torch.cuda.synchronize()
        dist.barrier()
        win.close()
    except sb.ScimlopError as e:
        out["replicated_fused_peer_unavailable"] = str(e)[:160]
    try:
        mcw = sd.MulticastWindow(h, (n, KCOLS), np.float64, rank, world, gather_obj, dist.barrier)
    except Exception as e:                      # no NVSwitch multicast on this box
        mcw = None
        out["replicated_fused_multicast_unavailable"] = str(e)[:160]
    if mcw is not None:
        out["replicated_fused_multicast_ms"] = _timed_max(torch, dist, world, lambda: sd.kron_mul_mc(h, L, V, mcw, kper), reps=20)
        h.set_option(sb.OPT_PEER_PIPELINE, 0)
        out["replicated_fused_multicast_unpipelined_ms"] = _timed_max(torch, dist, world, lambda: sd.kron_mul_mc(h, L, V, mcw, kper), reps=20)
        h.set_option(sb.OPT_PEER_PIPELINE, 1)
        mcw.W.t.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        sd.kron_mul_mc(h, L, V, mcw, kper)
        torch.cuda.synchronize()
        dist.barrier()
        good = check(mcw.W, "fused multicast all-gather")
        out["replicated_fused_multicast_bit_identical"] = good
        ok = ok and good
        torch.cuda.synchronize()
        dist.barrier()
        mcw.close()
    flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    out["all_ranks_bit_identical"] = bool(flag.item() == 0)
    return out


def c3_leg(torch, dist, sb, world, rank):
    """configs[2]: kron(A, B, C) with 128 x 128 Float32 factors on a 2097152 x 512 batch per GPU (column-sharded like the
    headline): modes (C, B) fused on chip by kron2_f16_kernel, mode A by the resident-factor tcgen05 kernel.  HBM-bound:
    the fraction is ALGORITHMIC bytes (V once + W once) over the measured copy bandwidth; the schedule moves 2.0x that."""
    n, k = 128 ** 3, 512
    g = torch.Generator(device="cuda").manual_seed(33 + rank)
    F = [sb.DeviceArray(torch.rand(128 * 128, dtype=torch.float32, device="cuda", generator=g) - 0.5, (128, 128)) for _ in range(3)]
    V = sb.DeviceArray(torch.rand(n * k, dtype=torch.float32, device="cuda", generator=g), (n, k))
    W = sb.DeviceArray.empty((n, k), np.float32)
    L = sb.cache_operator(sb.TensorProductOperator(*[sb.MatrixOperator(f) for f in F]), V)
    ms = _timed_max(torch, dist, world, lambda: sb.mul_(W, L, V), reps=5, warm=2)
    # parity spot check: one column against the Float64 product of the three modes, on the device
    j = 7
    X = V.cols(j, j + 1).t.double().reshape(128, 128, 128)               # [a][b][c]
    Fd = [f.t.double().reshape(128, 128).t() for f in F]                  # row-major views of the column-major factors
    Y = X @ Fd[2].t()                                                     # mode C: [a][b][c'] 
    Y = (Y.transpose(1, 2) @ Fd[1].t()).transpose(1, 2)                  # mode B: [a][b'][c']
    ref = torch.tensordot(Fd[0], Y, dims=([1], [0])).reshape(-1)         # mode A: [a'][b'][c']
    got = W.cols(j, j + 1).t.double()
    err = float((got - ref).norm() / ref.norm())
    peak = hbm_peak_gbs()
    gbs = 2.0 * n * k * 4 / (ms * 1e-3) / 1e9
    out = {"workload": "configs[2]: TensorProductOperator of three 128x128 Float32 MatrixOperators, 2097152 x 512 batch per GPU",
           "ms": ms, "tflops_fp32_equiv_aggregate": 3 * 2 * 128.0 * n * k * world / (ms * 1e-3) / 1e12,
           "algorithmic_gbs_per_gpu": gbs, "frac_of_hbm_copy_peak": gbs / peak[0], "hbm_peak": peak[1],
           "passes": 2, "parity_rel_err_vs_f64_column": err}
    del V, W, L
    torch.cuda.empty_cache()
    return out


def hbm_peak_gbs():
    try:
        mp = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))
        return float(mp["hbm_gbs"]), "MEASURED_PEAKS.json (copy bandwidth measured on this pool)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def c5_leg(torch, dist, sb, world, rank):
    """configs[4]: kron(Dxx, I) + kron(I, Dyy) on a 4096 x 4096 grid (dense spectral-style factors, Float64), batch-
    sharded: ONE column per GPU, C5_MATVECS dependent matvecs (w -> v ping-pong) captured in one CUDA graph."""
    N = C5_N
    rng = np.random.default_rng(5)
    Dxx = np.asfortranarray(rng.random((N, N)) / N)      # spectral radius ~ 0.5: repeated application stays finite
    Dyy = np.asfortranarray(rng.random((N, N)) / N)
    g = torch.Generator(device="cuda").manual_seed(77 + rank)
    v = sb.DeviceArray(torch.rand(N * N, dtype=torch.float64, device="cuda", generator=g), (N * N, 1))
    w = sb.DeviceArray.zeros((N * N, 1), np.float64)
    Ix = sb.IdentityOperator(N)
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(Dxx), Ix) + sb.kron(Ix, sb.MatrixOperator(Dyy)), v)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        sb.mul_(w, L, v)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for i in range(C5_MATVECS // 2):
                sb.mul_(w, L, v)
                sb.mul_(v, L, w)
    torch.cuda.current_stream().wait_stream(side)
    ms = _timed_max(torch, dist, world, graph.replay, reps=2, warm=1)
    flops_col = 2.0 * 2.0 * N ** 3
    finite = bool(torch.isfinite(v.t).all().item())
    return {"workload": "configs[4]: AddedOperator kron(Dxx,I)+kron(I,Dyy), dense 4096x4096 Float64 factors, "
                        "N = 16777216, one column per GPU (global batch = n_gpus columns)",
            "matvecs_per_graph": C5_MATVECS, "ms_per_matvec": ms / C5_MATVECS,
            "tflops_aggregate": flops_col * world / (ms / C5_MATVECS * 1e-3) / 1e12,
            "frac_of_dmma_peak_per_gpu": flops_col / (ms / C5_MATVECS * 1e-3) / 1e12 / FP64_DMMA_PEAK_TFLOPS,
            "result_finite": finite}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import scimlop_b200 as sb
    h = sb.handle(local)

    A, B = make_factors()
    n = NO * NI
    # every rank owns its own 256-column slab of the global batch (seeded per rank)
    Vh = torch.empty(n * KCOLS, dtype=torch.float64).pin_memory()
    Vh.numpy()[:] = np.random.default_rng(1 + 1000 * rank).random(n * KCOLS)
    Wh = torch.zeros(MO * MI * KCOLS, dtype=torch.float64).pin_memory()
    V = sb.DeviceArray(Vh.to("cuda", non_blocking=False), (n, KCOLS))
    W = sb.DeviceArray.full((MO * MI, KCOLS), float("nan"), np.float64)
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)), V)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------
    for _ in range(args.warmup):
        sb.mul_(W, L, V)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = h.launch_count()
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            sb.mul_(W, L, V)
        e1.record()
        e1.synchronize()
    barrier()
    launches = h.launch_count() - l0
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    flops_step_rank = FLOPS_PER_COL * KCOLS
    value = flops_step_rank * world / (ms * 1e-3) / 1e12

    if args.kernel_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": ms, "kernel_only": True,
                              "gpu_launches": int(launches), "clocks": clk.summary()}))
        if world > 1:
            dist.destroy_process_group()
        return 0

    # 5-arg form (α, β) for the record
    W.t.fill_(0.5)
    for _ in range(2):
        sb.mul_(W, L, V, 0.7, 0.3)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(max(3, args.steps // 4)):
        sb.mul_(W, L, V, 0.7, 0.3)
    e1.record()
    e1.synchronize()
    ms5 = e0.elapsed_time(e1) / max(3, args.steps // 4)

    # ---- end to end through the host-buffer C-ABI call (pinned host V/W, copies inside the timed region) ----
    Ah, Bh = np.asfortranarray(A), np.asfortranarray(B)
    ptrs = (C.c_void_p * 2)(Ah.ctypes.data, Bh.ctypes.data)
    dims = (C.c_int64 * 4)(MO, NO, MI, NI)
    lds = (C.c_int64 * 2)(MO, MI)
    kinds = (C.c_int32 * 2)(0, 0)

    def e2e_call():
        h.check(h.lib.scimlop_kron_mul_host(h.ptr, 0, 0, 2, ptrs, dims, lds, kinds, C.c_void_p(Vh.data_ptr()), n,
                                            C.c_void_p(Wh.data_ptr()), MO * MI, KCOLS, None, None, 0),
                "scimlop_kron_mul_host")

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = flops_step_rank * world / e2e_s / 1e12
    # the e2e result must be the same numbers as the device-resident path
    sb.mul_(W, L, V)
    torch.cuda.synchronize()
    mism = float((W.t[: MO * MI * 4].cpu() - Wh[: MO * MI * 4]).abs().max())
    W_head = W.cols(0, 2).numpy()          # for the parity spot check below (W is reused as a copy target next)

    # pure-copy floor of the e2e step (H2D of V concurrent with D2H of W, all ranks at once)
    floor_ms = copy_floor_ms(torch, dist, world, Vh, Wh, V.t, W.t)

    # ---- strong-scaled named shape (+ replicated-W variants) and configs[4]; both inside `config` -------------
    del L
    strong = strong_legs(torch, dist, sb, h, world, rank, A, B, ms)
    c5 = c5_leg(torch, dist, sb, world, rank)
    c3 = c3_leg(torch, dist, sb, world, rank)
    if world > 1 and not strong.get("all_ranks_bit_identical", True):
        if rank == 0:
            print(json.dumps({"error": "a fused compute + all-gather variant differs from the unfused result", "strong": strong}))
        dist.destroy_process_group()
        return 3

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ----------------------------------------
    # stdlib variant on the whole batch (same configuration as the GPU arm, ~1.3 s per mul!); the LoopVectorization
    # variant (single-thread SIMD nest) on its own 32-column sample, reported in its own TFLOP/s -- nothing is scaled
    vs = np.asfortranarray(Vh.numpy().reshape((n, KCOLS), order="F"))
    cpu_line = None
    if world == 1:
        try:
            from threadpoolctl import threadpool_limits
            limiter = threadpool_limits(limits=os.cpu_count())
        except Exception:
            limiter = None
        t_std = cpu_time_one("stdlib", A, B, vs, reps=2)
        t_lv = cpu_time_one("lv", A, B, vs[:, :LV_SAMPLE_COLS].copy(order="F"), reps=2)
        std_tf = FLOPS_PER_COL * KCOLS / t_std / 1e12
        lv_tf = FLOPS_PER_COL * LV_SAMPLE_COLS / t_lv / 1e12
        cpu_line = {"value": max(std_tf, lv_tf), "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "port",
                    "sample": f"stdlib variant: all {KCOLS} columns (min of 2 runs after one warm-up); LoopVectorization "
                              f"variant: {LV_SAMPLE_COLS} columns, unscaled",
                    "stdlib_path_s": t_std, "stdlib_tflops": std_tf, "lv_path_s": t_lv, "lv_tflops": lv_tf,
                    "lv_cores": 1, "note": "CPU restatement of the reference algorithm (oracle/), not the Julia reference itself"}
    # parity spot check of the timed GPU result against the same CPU restatement
    from oracle import ref_numpy as R
    err = R.rel_err(W_head, R.kron_apply([A, B], np.asfortranarray(vs[:, :2])))

    flops_per_launch = flops_step_rank / 2.0           # both launches are the same kernel and the same flop count
    launch_ms = ms / 2.0
    achieved = flops_per_launch / (launch_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: TensorProductOperator(MatrixOperator(A), MatrixOperator(B)), A,B 512x512 "
                               "Float64, cache_operator + mul!(w,L,v) on a 262144x256 batch per GPU",
                   "global_batch_columns": KCOLS * world, "parallelism": f"column-sharded x{world}, factors replicated, "
                   "no data-path collective", "l2": "inputs larger than L2: V, W and the intermediate are 537 MB "
                   "each vs 126 MB L2", "ms_per_step_5arg": ms5, "parity_rel_err_vs_oracle": err,
                   "e2e_vs_resident_max_abs_diff": mism, "strong": strong, "c5": c5, "c3": c3},
        "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": "TFLOP/s", "h2d_bytes_per_step": int(Vh.numel() * 8 + Ah.nbytes + Bh.nbytes),
                "d2h_bytes_per_step": int(Wh.numel() * 8), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                "copy_floor_ms": floor_ms, "frac_of_copy_floor": floor_ms / (e2e_s * 1e3),
                "copy_floor": "pinned H2D of V concurrent with D2H of W on two streams, no kernels, all ranks at once "
                              "(max over ranks); the ranks of one box share one NUMA node and its PCIe / memory paths",
                "api": "scimlop_kron_mul_host (pinned host V/W, chunked H2D | kernels | D2H overlap)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                     "frac": achieved / FP64_DMMA_PEAK_TFLOPS, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                     "traffic_unit": "bytes of DRAM traffic per launch (ncu --set full, profiles/r02_ncu_dmma_summary.txt; round 1: r01_ncu_dmma_v6_summary.txt); "
                                     "algorithmic bytes per launch = 1.076e9",
                     "kernel": "scimlop::dmma::gemm_dmma_kernel (2 launches per step: stage 1 B*U, stage 2 batched C1_j*A^T)",
                     "peak_source": "FP64 DMMA.8x8x4 issue-rate microbenchmark on this pool (profiles/r01_peaks_microbench.json); "
                                    "MEASURED_PEAKS.json holds no FP64 figure (cuBLAS DGEMM 8192^3 reaches 35.5)",
                     "algorithmic_flops_per_launch": flops_per_launch, "launch_ms": launch_ms},
        "cpu_baseline": cpu_line,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kernel-only", action="store_true",
                    help="device-resident timing only (for ncu runs): skips the e2e and cpu_baseline legs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
