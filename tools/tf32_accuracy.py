"""Error of the Float32 GEMM paths against Float64 as a function of the contraction length, for all-positive and
for centred data (tensor-core FP32 accumulation rounds toward zero: positive data shows the bias)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
from scimlop_b200 import _lib
rng = np.random.default_rng(0)
M, N = 512, 512
for K in (128, 512, 2048, 8192, 32768):
    for name, shift in (("positive", 0.0), ("centred", 0.5)):
        A = np.asfortranarray((rng.random((M, K)) - shift).astype(np.float32))
        B = np.asfortranarray((rng.random((K, N)) - shift).astype(np.float32))
        want = A.astype(np.float64) @ B.astype(np.float64)
        scale = np.linalg.norm(np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64))
        out = {"K": K, "data": name}
        for prec, pn in ((_lib.PREC_DEFAULT, "tf32x3"), (_lib.PREC_FP32_EXACT, "ffma")):
            L = sb.MatrixOperator(A); L.precision = prec
            w = sb.DeviceArray.zeros((M, N), np.float32)
            sb.mul_(w, L, sb.DeviceArray.from_numpy(B))
            out[pn] = float(np.linalg.norm(w.numpy().astype(np.float64) - want) / scale)
        print(json.dumps(out), flush=True)
