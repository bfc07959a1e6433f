"""Per-role timeline of the fused two-mode kernel (csrc/kron2_f16.cuh): builds a -DSCIMLOP_KF16_TRACE copy of the library
(`--build`, run here: nvcc cross-compiles; the .so travels to the GPU box), runs one 128 (x) 128 Float32 product and
prints, for steady-state slabs of CTA 0, the clock64() offsets of every pipeline event relative to the slab's first MMA.

    python tools/kf16_trace.py --build            # in the container
    gpurun -- python tools/kf16_trace.py          # on the B200
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, "scimloperators.jl_b200", "libscimlop_b200_trace.so")

SLOTS = {0: "prod kk0 slot free", 1: "prod kk1", 2: "prod kk2", 3: "prod kk3", 8: "mma S1 issue start (x_full seen)", 9: "mma S1 issued",
         15: "mma S2 issue start (y_full seen)", 16: "mma S2 issued",
         20: "spl raw_full0", 21: "spl raw_full1", 22: "spl raw_full2", 23: "spl raw_full3", 28: "spl g0 scale known", 29: "spl g1 scale known",
         32: "spl g0 x_free seen", 33: "spl g1 x_free seen",
         52: "spl g0 q0 done", 53: "spl g0 q1 done", 54: "spl g0 q2 done", 55: "spl g0 q3 done", 56: "spl g1 q0 done", 57: "spl g1 q1 done",
         58: "spl g1 q2 done", 59: "spl g1 q3 done", 40: "drain1 start", 41: "drain1 done", 44: "drain2 start", 46: "drain2 accumulator handed back", 45: "drain2 done"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--k", type=int, default=148 * 48)
    ap.add_argument("--beta", type=float, default=0.0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--exp", type=int, default=0, help="timing experiment compiled in with -DKF16_EXP=n (results are wrong)")
    a = ap.parse_args()
    if a.build:
        import importlib.util
        spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "scimloperators.jl_b200", "build.py"))
        b = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(b)
        out = TRACE_LIB.replace(".so", f"_exp{a.exp}.so") if a.exp else TRACE_LIB
        print(b.build(force=True, out=out, extra=["-DSCIMLOP_KF16_TRACE"] + ([f"-DKF16_EXP={a.exp}"] if a.exp else [])))
        return
    import numpy as np
    import torch
    import scimlop_b200 as sb
    from scimlop_b200 import _lib
    _lib.LIB_PATH = TRACE_LIB.replace(".so", f"_exp{a.exp}.so") if a.exp else TRACE_LIB
    h = sb.handle(0)
    trace = torch.zeros(64 * 64, dtype=torch.int64, device="cuda")
    h.lib.scimlop_debug_kf16_trace.argtypes = [C.c_void_p]
    h.lib.scimlop_debug_kf16_trace(C.c_void_p(trace.data_ptr()))
    g = torch.Generator(device="cuda").manual_seed(0)
    F = [sb.DeviceArray(torch.rand(128 * 128, dtype=torch.float32, device="cuda", generator=g), (128, 128)) for _ in range(2)]
    V = sb.DeviceArray(torch.rand(128 * 128 * a.k, dtype=torch.float32, device="cuda", generator=g), (128 * 128, a.k))
    W = sb.DeviceArray.full((128 * 128, a.k), 0.5, np.float32)
    L = sb.cache_operator(sb.TensorProductOperator(*[sb.MatrixOperator(f) for f in F]), V)
    for _ in range(a.reps):
        if a.beta:
            sb.mul_(W, L, V, 1.0, a.beta)
        else:
            sb.mul_(W, L, V)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        if a.beta:
            sb.mul_(W, L, V, 1.0, a.beta)
        else:
            sb.mul_(W, L, V)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"ms_per_launch": ms, "us_per_slab_per_sm": ms * 1e3 / (a.k / 148.0)}))
    t = trace.cpu().numpy().reshape(64, 64)
    period = np.diff(t[8:40, 15])
    ns = np.diff(t[8:40, 51])
    print(json.dumps({"k": a.k, "slab_period_clk_median": float(np.median(period)), "min": int(period.min()), "max": int(period.max()),
                      "slab_period_ns_median": float(np.median(ns)), "sm_ghz": float(np.sum(period) / max(1, np.sum(ns)))}))
    for it in (20, 21):
        base = t[it, 15]
        ev = sorted((int(t[it, s] - base), name) for s, name in SLOTS.items() if t[it, s])
        print(f"--- slab iteration {it} (clk relative to the slab's stage-2 issue; next slab's at {int(t[it + 1, 15] - base)})")
        for d, name in ev:
            print(f"{d:8d}  {name}")


if __name__ == "__main__":
    main()
