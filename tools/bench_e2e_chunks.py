"""End-to-end (host buffers) time of configs[1] through scimlop_kron_mul_host for several chunk schedules, next to the
pure-copy floor (pinned H2D of V concurrent with D2H of W, no kernels): where the distance to the floor comes from.

    gpurun -- python tools/bench_e2e_chunks.py > profiles/r02_e2e_chunks.jsonl
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import scimlop_b200 as sb
    h = sb.handle(0)
    MO = NO = MI = NI = 512
    K = 256
    n = NO * NI
    rng = np.random.default_rng(0)
    A, B = np.asfortranarray(rng.random((MO, NO))), np.asfortranarray(rng.random((MI, NI)))
    Vh = torch.rand(n * K, dtype=torch.float64).pin_memory()
    Wh = torch.zeros(MO * MI * K, dtype=torch.float64).pin_memory()
    ptrs = (C.c_void_p * 2)(A.ctypes.data, B.ctypes.data)
    dims = (C.c_int64 * 4)(MO, NO, MI, NI)
    lds = (C.c_int64 * 2)(MO, MI)
    kinds = (C.c_int32 * 2)(0, 0)

    def call(chunk):
        h.check(h.lib.scimlop_kron_mul_host(h.ptr, 0, 0, 2, ptrs, dims, lds, kinds, C.c_void_p(Vh.data_ptr()), n,
                                            C.c_void_p(Wh.data_ptr()), MO * MI, K, None, None, chunk), "scimlop_kron_mul_host")

    def timed(fn, reps=8):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        return min(ts), sorted(ts)[len(ts) // 2]

    Vd = torch.empty(n * K, dtype=torch.float64, device="cuda")
    Wd = torch.empty(MO * MI * K, dtype=torch.float64, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def floor():
        with torch.cuda.stream(s_up):
            Vd.copy_(Vh, non_blocking=True)
        with torch.cuda.stream(s_dn):
            Wh.copy_(Wd, non_blocking=True)

    def up_only():
        with torch.cuda.stream(s_up):
            Vd.copy_(Vh, non_blocking=True)

    def dn_only():
        with torch.cuda.stream(s_dn):
            Wh.copy_(Wd, non_blocking=True)

    for name, fn in (("copy floor: H2D of V || D2H of W", floor), ("H2D of V alone", up_only), ("D2H of W alone", dn_only)):
        tmin, tmed = timed(fn)
        print(json.dumps({"what": name, "ms_min": round(tmin, 3), "ms_median": round(tmed, 3),
                          "gbs_per_direction": round(n * K * 8 / (tmed * 1e-3) / 1e9, 1)}), flush=True)
    for chunk in (0, 2, 4, 8, 16, 32, 64, 128):
        tmin, tmed = timed(lambda: call(chunk))
        print(json.dumps({"what": "scimlop_kron_mul_host", "chunk_cols": chunk if chunk else "auto (ramped, 16)", "ms_min": round(tmin, 3),
                          "ms_median": round(tmed, 3)}), flush=True)


if __name__ == "__main__":
    main()
