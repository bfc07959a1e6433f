"""Phase timeline of scimlop_factorize (csrc/lu.cu, look-ahead LU): builds a -DSCIMLOP_LU_TRACE copy of the library
(`--build`, in the container) and factorises one n x n matrix with it on the B200; the library prints, for a sample of
panel steps, when the panel kernel ran on the caller's stream and when the side stream's swaps / trailing update ran.

    python tools/lu_trace.py --build
    gpurun -- python tools/lu_trace.py 4096
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, "scimloperators.jl_b200", "libscimlop_b200_lutrace.so")

if "--build" in sys.argv:
    import importlib.util
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "scimloperators.jl_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    print(b.build(force=True, out=TRACE_LIB, extra=["-DSCIMLOP_LU_TRACE"]))
    sys.exit(0)
import torch
import scimlop_b200 as sb
from scimlop_b200 import _lib
_lib.LIB_PATH = TRACE_LIB
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator(device="cuda").manual_seed(0)
A = sb.DeviceArray(torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g) - 0.5, (n, n))
lu = sb.LUFactorization(A)
torch.cuda.synchronize()
print("---- second factorisation (warm) ----", file=sys.stderr, flush=True)
lu.refresh()
torch.cuda.synchronize()
print("ok", lu.cond)
