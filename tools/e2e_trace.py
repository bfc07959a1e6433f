"""Per-chunk timeline of scimlop_kron_mul_host (csrc/extras.cu) on configs[1]: builds a -DSCIMLOP_HOST_TRACE copy of the
library (`--build`, in the container); on the B200 the library prints, for every chunk, when its upload, its two GEMM
launches and its download ran (device time) and when the host had finished enqueueing it.

    python tools/e2e_trace.py --build
    gpurun -- python tools/e2e_trace.py
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NOCOMPUTE = "--nocompute" in sys.argv          # timing experiment: chunked copies only (results are garbage)
TRACE_LIB = os.path.join(ROOT, "scimloperators.jl_b200", "libscimlop_b200_hosttrace_nc.so" if NOCOMPUTE else "libscimlop_b200_hosttrace.so")

if "--build" in sys.argv:
    import importlib.util
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "scimloperators.jl_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    print(b.build(force=True, out=TRACE_LIB, extra=["-DSCIMLOP_HOST_TRACE"] + (["-DSCIMLOP_HOST_TRACE_NOCOMPUTE"] if NOCOMPUTE else [])))
    sys.exit(0)
import numpy as np
import torch
import scimlop_b200 as sb
from scimlop_b200 import _lib
_lib.LIB_PATH = TRACE_LIB
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
nums = [a for a in sys.argv[1:] if a.lstrip("-").isdigit()]
chunk = int(nums[0]) if nums else 0
for rep in range(3):
    print(f"---- call {rep} ----", file=sys.stderr, flush=True)
    h.check(h.lib.scimlop_kron_mul_host(h.ptr, 0, 0, 2, ptrs, dims, lds, kinds, C.c_void_p(Vh.data_ptr()), n,
                                        C.c_void_p(Wh.data_ptr()), MO * MI, K, None, None, chunk), "scimlop_kron_mul_host")
torch.cuda.synchronize()
import time
ts = []
for _ in range(6):
    t0 = time.perf_counter()
    h.check(h.lib.scimlop_kron_mul_host(h.ptr, 0, 0, 2, ptrs, dims, lds, kinds, C.c_void_p(Vh.data_ptr()), n,
                                        C.c_void_p(Wh.data_ptr()), MO * MI, K, None, None, chunk), "scimlop_kron_mul_host")
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("ok; wall ms per call (with the trace events):", [round(t, 2) for t in ts])
