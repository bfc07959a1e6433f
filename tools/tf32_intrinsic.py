"""Round-2 probe: per-tile time of the tcgen05 resident-factor kernel when HBM is out of the picture.

Batches whose operand windows overlap (batch stride = 4 floats) keep every X tile and every D tile inside a few hundred
KB, so TMA loads and bulk stores hit L2: what remains is the kernel's own pipeline (TMA -> splitters -> TMEM -> UMMA ->
epilogue).  The result bounds what an on-chip fusion of two mode products can reach per 128 x 128 slab.
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb  # noqa: E402
from scimlop_b200 import _lib  # noqa: E402


def main():
    h = _lib.handle(0)
    lib = h.lib
    F = torch.rand(128 * 128, dtype=torch.float32, device="cuda") / 64
    for prec, name in ((_lib.PREC_DEFAULT, "3xTF32"), (_lib.PREC_TF32, "1xTF32")):
        for tiles_per_sm in (8, 32, 128):
            batch = 148 * tiles_per_sm
            X = torch.rand(128 * 128 + 4 * batch + 64, dtype=torch.float32, device="cuda")
            D = torch.empty_like(X)
            one = C.c_float(1.0)

            def run():
                rc = lib.scimlop_gemm_strided_batched(h.ptr, _lib.F32, prec, 0, 1, 128, 128, 128, C.byref(one), X.data_ptr(), 128, 4,
                                                      F.data_ptr(), 128, 0, None, D.data_ptr(), 128, 4, batch,
                                                      torch.cuda.current_stream().cuda_stream)
                h.check(rc, "gemm")

            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = float(np.median(ts))
            print(json.dumps({"probe": "tf32 resident, L2-fed overlapping tiles", "mode": name, "tiles_per_sm": tiles_per_sm,
                              "ms": round(t, 4), "us_per_tile": round(t * 1e3 / tiles_per_sm, 3),
                              "TFLOPs_fp32_equiv": round(2 * 128 ** 3 * batch / t / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
